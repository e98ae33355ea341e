// CPU emulation of a CUDA launch, for tests only (never part of the product library).
//
// Each CUDA thread of a CTA becomes a ucontext fiber on one OS thread; ctx.sync() yields to a
// round-robin scheduler, so one scheduler sweep is exactly one __syncthreads() phase.  CTAs run
// one after another.  This executes the very kernel bodies the device build compiles
// (csrc/abbe_kernels.h) and lets the index arithmetic be checked against the oracle without a GPU.
#pragma once
#include <ucontext.h>

#include "hd.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

namespace litho_emu {

struct Fiber {
    ucontext_t uc;
    char* stack = nullptr;
    bool done = false;
};

// A barrier with real arrival counting: a thread that arrives waits (yielding) until all `n` participants of
// this generation have arrived, so threads may yield any number of times elsewhere (mbarrier waits) without
// slipping through a barrier, and mismatched barriers deadlock visibly instead of passing silently.
struct Bar {
    int arrived = 0;
    unsigned gen = 0;
};

struct Sched {
    ucontext_t main_uc;
    std::vector<Fiber> fibers;
    int current = -1;
    std::function<void(int)> body;  // body(tid)
    Bar cta_bar[16];                // id 0 = __syncthreads, 1..15 named barriers
    std::vector<Bar> warp_bar;      // __syncwarp, one per warp
    long progress = 0;              // bumped whenever a barrier or an mbarrier phase completes
};

inline Sched*& cur_sched() {
    static thread_local Sched* s = nullptr;
    return s;
}

inline void fiber_entry() {
    Sched* s = cur_sched();
    int me = s->current;
    s->body(me);
    s->fibers[me].done = true;
    swapcontext(&s->fibers[me].uc, &s->main_uc);
}

inline void fiber_yield() {
    Sched* s = cur_sched();
    int me = s->current;
    swapcontext(&s->fibers[me].uc, &s->main_uc);
}

inline void barrier_wait(Bar& b, int n) {
    Sched* s = cur_sched();
    const unsigned g = b.gen;
    if (++b.arrived >= n) {
        b.arrived = 0;
        ++b.gen;
        ++s->progress;
    } else {
        while (b.gen == g) fiber_yield();
    }
}

struct EmuCtx {
    int tid_, bdim_, bx_, by_, bz_, gdx_;
    int tid() const { return tid_; }
    int bdim() const { return bdim_; }
    int bx() const { return bx_; }
    int by() const { return by_; }
    int bz() const { return bz_; }
    int gdx() const { return gdx_; }
    void sync() const { barrier_wait(cur_sched()->cta_bar[0], bdim_); }
    void sync_warp() const {
        const int w = tid_ / 32;
        const int n = (bdim_ - w * 32) < 32 ? (bdim_ - w * 32) : 32;
        barrier_wait(cur_sched()->warp_bar[w], n);
    }
    void sync_named(int id, int n) const { barrier_wait(cur_sched()->cta_bar[id & 15], n); }
    // asynchronous copies complete immediately in the emulation
    void cp_async16(void* dst, const void* src) const { memcpy(dst, src, 16); }
    void cp_async_wait() const {}
    // TMA tile copies complete at issue in the emulation (boxes past the tensor's last row are zero-filled,
    // like the hardware's out-of-bounds fill).  The mbarrier word counts completed phases; a waiter yields
    // until the phase of its parity has completed, as the hardware's try_wait.parity loop does.
    void mbar_init(unsigned long long* bar, int) const { *bar = 0; }
    void tile_load(void* dst, const litho::TileMap& tm, int row, int col, int nbox, int rim_rows, int rim_elem,
                   unsigned long long* bar) const {
        char* d = (char*)dst;
        for (int i = 0; i < nbox; ++i)
            for (int r = 0; r < tm.box_rows; ++r, d += (size_t)tm.box_cols * 8) {
                const long long y = row + (long long)i * tm.box_rows + r;
                if (y < tm.rows) memcpy(d, tm.base + y * tm.pitch + col, (size_t)tm.box_cols * 8);
                else memset(d, 0, (size_t)tm.box_cols * 8);
            }
        for (int k = 0; k < rim_rows; ++k)
            memcpy((char*)dst + (size_t)(rim_elem + k * tm.box_cols) * 8,
                   tm.base + (long long)(row + nbox * tm.box_rows + k) * tm.pitch + col, (size_t)tm.box_cols * 8);
        ++*bar;
        ++cur_sched()->progress;
    }
    bool mbar_wait(unsigned long long* bar, unsigned parity) const {
        while (!(*bar > 0 && ((*bar - 1) & 1) == parity)) fiber_yield();
        return true;
    }
};

// run one CTA of `nthreads` fibers; body(ctx) is the kernel body bound to its parameters
template <class Body>
void run_cta(int nthreads, int bx, int by, int bz, int gdx, Body body) {
    static const size_t STACK = 256 * 1024;
    Sched s;
    s.fibers.resize(nthreads);
    s.body = [&](int tid) {
        EmuCtx ctx{tid, nthreads, bx, by, bz, gdx};
        body(ctx);
    };
    s.warp_bar.resize((nthreads + 31) / 32);
    cur_sched() = &s;
    for (int t = 0; t < nthreads; ++t) {
        Fiber& f = s.fibers[t];
        f.stack = (char*)malloc(STACK);
        getcontext(&f.uc);
        f.uc.uc_stack.ss_sp = f.stack;
        f.uc.uc_stack.ss_size = STACK;
        f.uc.uc_link = &s.main_uc;
        makecontext(&f.uc, (void (*)())fiber_entry, 0);
    }
    bool any = true;
    int idle_sweeps = 0;
    static const int order = []() {
        const char* e = getenv("LITHO_EMU_ORDER");
        return !e ? 0 : (!strcmp(e, "reverse") ? 1 : (!strcmp(e, "random") ? 2 : 0));
    }();
    // a stride coprime with the thread count visits every thread once per sweep
    unsigned stride = 1;
    if (order == 2) {
        auto gcd = [](unsigned a, unsigned b) { while (b) { unsigned r = a % b; a = b; b = r; } return a; };
        for (stride = (unsigned)nthreads / 2 + 1; gcd(stride, (unsigned)nthreads) != 1; ++stride) {}
    }
    unsigned sweep = 0;
    while (any) {
        any = false;
        const long before = s.progress;
        int finished = 0;
        // Order in which the threads of the CTA run between two barriers: ascending by default; LITHO_EMU_ORDER =
        // reverse | random makes a missing barrier (a result that depends on who runs first) show up as a
        // wrong answer in the parity tests.
        ++sweep;
        for (int i = 0; i < nthreads; ++i) {
            int t = i;
            if (order == 1) t = nthreads - 1 - i;
            else if (order == 2) t = (int)(((unsigned long long)i * stride + (unsigned long long)sweep * 7919u) % (unsigned)nthreads);
            if (s.fibers[t].done) continue;
            s.current = t;
            swapcontext(&s.main_uc, &s.fibers[t].uc);
            if (!s.fibers[t].done) any = true;
            else ++finished;
        }
        // a sweep in which no barrier completed and no thread finished means the CTA is deadlocked
        idle_sweeps = (s.progress == before && finished == 0) ? idle_sweeps + 1 : 0;
        if (idle_sweeps > 4) {
            fprintf(stderr, "litho_emu: CTA (%d,%d,%d) deadlocked (mismatched barrier or mbarrier never completed)\n", bx, by, bz);
            abort();
        }
    }
    for (auto& f : s.fibers) free(f.stack);
    cur_sched() = nullptr;
}

template <class Body>
void launch(int gx, int gy, int gz, int nthreads, size_t smem_bytes, Body body) {
    std::vector<char> smem(smem_bytes + 64);
    for (int z = 0; z < gz; ++z)
        for (int y = 0; y < gy; ++y)
            for (int x = 0; x < gx; ++x) {
                memset(smem.data(), 0xCD, smem.size());  // poison: uninitialised reads show up as garbage
                run_cta(nthreads, x, y, z, gx, [&](const EmuCtx& ctx) { body(ctx, smem.data()); });
            }
}

}  // namespace litho_emu
