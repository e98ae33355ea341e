// CPU emulation of a CUDA launch, for tests only (never part of the product library).
//
// Each CUDA thread of a CTA becomes a ucontext fiber on one OS thread; ctx.sync() yields to a
// round-robin scheduler, so one scheduler sweep is exactly one __syncthreads() phase.  CTAs run
// one after another.  This executes the very kernel bodies the device build compiles
// (csrc/abbe_kernels.h) and lets the index arithmetic be checked against the oracle without a GPU.
#pragma once
#include <ucontext.h>

#include "hd.h"

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

struct Sched {
    ucontext_t main_uc;
    std::vector<Fiber> fibers;
    int current = -1;
    std::function<void(int)> body;  // body(tid)
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

struct EmuCtx {
    int tid_, bdim_, bx_, by_, bz_, gdx_;
    int tid() const { return tid_; }
    int bdim() const { return bdim_; }
    int bx() const { return bx_; }
    int by() const { return by_; }
    int bz() const { return bz_; }
    int gdx() const { return gdx_; }
    void sync() const { fiber_yield(); }
    // a warp barrier is emulated by the (stronger) CTA-wide phase boundary; every thread of the CTA
    // executes the same number of them in the kernels that use it
    void sync_warp() const { fiber_yield(); }
    void sync_named(int, int) const { fiber_yield(); }
    // asynchronous copies complete immediately in the emulation
    void cp_async16(void* dst, const void* src) const { memcpy(dst, src, 16); }
    void cp_async_wait() const {}
    // TMA tile copies complete at issue in the emulation (boxes past the tensor's last row are zero-filled,
    // like the hardware's out-of-bounds fill); the barrier is then always satisfied
    void mbar_init(unsigned long long*, int) const {}
    void tile_load(void* dst, const litho::TileMap& tm, int row, int col, int nbox, unsigned long long*) const {
        char* d = (char*)dst;
        for (int i = 0; i < nbox; ++i)
            for (int r = 0; r < tm.box_rows; ++r, d += (size_t)tm.box_cols * 8) {
                const long long y = row + (long long)i * tm.box_rows + r;
                if (y < tm.rows) memcpy(d, tm.base + y * tm.pitch + col, (size_t)tm.box_cols * 8);
                else memset(d, 0, (size_t)tm.box_cols * 8);
            }
    }
    void mbar_wait(unsigned long long*, unsigned) const {}
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
    while (any) {
        any = false;
        for (int t = 0; t < nthreads; ++t) {
            if (s.fibers[t].done) continue;
            s.current = t;
            swapcontext(&s.main_uc, &s.fibers[t].uc);
            if (!s.fibers[t].done) any = true;
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
