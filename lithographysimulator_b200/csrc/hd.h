// Host/device-portable primitives for the Abbe imaging kernels.
//
// Every kernel body in this directory is written against a small "thread context"
// (tid / block ids / sync) so that the very same source compiles (a) as sm_100a device
// code and (b) as plain C++ that tests/emu runs thread-for-thread on the CPU to check
// index arithmetic without a GPU.  The CPU build is test infrastructure only; the
// product library contains the device build alone.
#pragma once
#include <math.h>
#include <stddef.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define LITHO_HD __host__ __device__ __forceinline__
#define LITHO_D __device__ __forceinline__
#else
#define LITHO_HD inline
#define LITHO_D inline
#endif
// Value-range hint for the device compiler (index arithmetic such as (g + TG*e)/PPT folds to constants once
// it knows g < TG); checked for real in the CPU emulation.
#if defined(__CUDA_ARCH__)
#define LITHO_ASSUME(x) __builtin_assume(x)
#elif defined(LITHO_EMU)
#include <assert.h>
#define LITHO_ASSUME(x) assert(x)
#else
#define LITHO_ASSUME(x) ((void)0)
#endif

namespace litho {

struct alignas(8) cplx {
    float x, y;
};

LITHO_HD cplx mk(float x, float y) { cplx c; c.x = x; c.y = y; return c; }
LITHO_HD cplx cadd(cplx a, cplx b) { return mk(a.x + b.x, a.y + b.y); }
LITHO_HD cplx csub(cplx a, cplx b) { return mk(a.x - b.x, a.y - b.y); }
LITHO_HD cplx cmul(cplx a, cplx b) { return mk(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
LITHO_HD cplx cconj(cplx a) { return mk(a.x, -a.y); }
LITHO_HD float cnorm2(cplx a) { return a.x * a.x + a.y * a.y; }
// multiply by +i / -i
LITHO_HD cplx mul_pi(cplx a) { return mk(-a.y, a.x); }
LITHO_HD cplx mul_mi(cplx a) { return mk(a.y, -a.x); }

struct int2_ {
    int x, y;
};

// read-only global load (LDG through the non-coherent path on the device)
LITHO_HD cplx ldg_c(const cplx* p) {
#if defined(__CUDA_ARCH__)
    float2 v = __ldg(reinterpret_cast<const float2*>(p));
    return mk(v.x, v.y);
#else
    return *p;
#endif
}
LITHO_HD float ldg_f(const float* p) {
#if defined(__CUDA_ARCH__)
    return __ldg(p);
#else
    return *p;
#endif
}

// L2-coherent (cache-global) accesses for data produced by other CTAs of the same launch
LITHO_HD cplx ldcg_c(const cplx* p) {
#if defined(__CUDA_ARCH__)
    float2 v = __ldcg(reinterpret_cast<const float2*>(p));
    return mk(v.x, v.y);
#else
    return *p;
#endif
}
LITHO_HD float ldcg_f(const float* p) {
#if defined(__CUDA_ARCH__)
    return __ldcg(p);
#else
    return *p;
#endif
}
LITHO_HD void stcg_f(float* p, float v) {
#if defined(__CUDA_ARCH__)
    __stcg(p, v);
#else
    *p = v;
#endif
}
// device-scope counter helpers
LITHO_HD int atomic_add_i(int* p, int v) {
#if defined(__CUDA_ARCH__)
    return atomicAdd(p, v);
#else
    int o = *p;
    *p += v;
    return o;
#endif
}
LITHO_HD int ld_volatile_i(const int* p) {
#if defined(__CUDA_ARCH__)
    return *reinterpret_cast<const volatile int*>(p);
#else
    return *p;
#endif
}
LITHO_HD void fence_gpu() {
#if defined(__CUDA_ARCH__)
    __threadfence();
#endif
}
LITHO_HD void backoff() {
#if defined(__CUDA_ARCH__)
    __nanosleep(200);
#endif
}

// floor-mod for a possibly negative a, b > 0
LITHO_HD int imod(int a, int b) {
    int m = a % b;
    return m < 0 ? m + b : m;
}
// ceil(a / b) for b > 0 and any sign of a
LITHO_HD int cdiv(int a, int b) { return a >= 0 ? (a + b - 1) / b : -((-a) / b); }
// floor(a / b) for b > 0 and any sign of a
LITHO_HD int fdiv(int a, int b) { return a >= 0 ? a / b : -((-a + b - 1) / b); }

// Column tile of the T ring as the TMA engine sees it: a 2-D tensor of float32 with 2*M floats per row
// (one complex64 row of T) and `rows` rows.  `map` is the CUtensorMap of the device build (opaque here, 64-byte
// aligned as the hardware requires); base/pitch/rows are the same facts in plain form, used by the CPU
// emulation of the copy (tests/emu) and for bounds.
struct alignas(64) TileMap {
    unsigned char map[128];
    const cplx* base;      // element (row 0, column 0)
    long long pitch;       // elements per row
    long long rows;        // rows of the tensor; boxes reaching past it are zero-filled
    int box_rows, box_cols;  // box of one copy: box_rows x box_cols complex elements
};

#if defined(__CUDACC__)
// Device thread context: thin wrapper over the CUDA built-ins.
struct DevCtx {
    __device__ __forceinline__ int tid() const { return threadIdx.x; }
    __device__ __forceinline__ int bdim() const { return blockDim.x; }
    __device__ __forceinline__ int bx() const { return blockIdx.x; }
    __device__ __forceinline__ int by() const { return blockIdx.y; }
    __device__ __forceinline__ int bz() const { return blockIdx.z; }
    __device__ __forceinline__ int gdx() const { return gridDim.x; }
    __device__ __forceinline__ void sync() const { __syncthreads(); }
    __device__ __forceinline__ void sync_warp() const { __syncwarp(); }
    // 16-byte asynchronous global -> shared copy (LDGSTS), and the wait for all of this thread's copies
    __device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) const {
        const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
    }
    __device__ __forceinline__ void cp_async_wait() const {
        asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
    }
    // named barrier over `n` threads (a multiple of 32); id 0 is __syncthreads' barrier, use 1..15
    __device__ __forceinline__ void sync_named(int id, int n) const {
        asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory");
    }
    // ---- TMA (cp.async.bulk.tensor) tile copies completing on a shared-memory mbarrier ----
    // One thread initialises the barrier (followed by a CTA barrier before anybody waits on it).
    __device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) const {
        const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(count) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // Issued by ONE thread: announce the bytes on the barrier, then copy `nbox` boxes of the tile whose first
    // element is (row, col) into consecutive box-sized pieces of `smem_dst` (128-byte aligned) and the
    // `rim_rows` rows that follow the boxes (row + nbox*box_rows + k) to elements rim_elem + k*box_cols of the
    // buffer with 1-D bulk copies.
    __device__ __forceinline__ void tile_load(void* smem_dst, const TileMap& tm, int row, int col, int nbox, int rim_rows,
                                              int rim_elem, unsigned long long* bar) const {
        const unsigned b = (unsigned)__cvta_generic_to_shared(bar);
        const unsigned box_bytes = (unsigned)(tm.box_rows * tm.box_cols) * 8u;
        const unsigned rim_bytes = (unsigned)tm.box_cols * 8u;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b),
                     "r"(box_bytes * (unsigned)nbox + rim_bytes * (unsigned)rim_rows)
                     : "memory");
        const unsigned d0 = (unsigned)__cvta_generic_to_shared(smem_dst);
        unsigned d = d0;
        const unsigned long long mp = reinterpret_cast<unsigned long long>(tm.map);
#pragma unroll 1
        for (int i = 0; i < nbox; ++i) {
            const int x = 2 * col;                       // float32 coordinates
            const int y = row + i * tm.box_rows;
            asm volatile(
                "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                ::"r"(d), "l"(mp), "r"(x), "r"(y), "r"(b)
                : "memory");
            d += box_bytes;
        }
#pragma unroll 1
        for (int k = 0; k < rim_rows; ++k) {
            const cplx* src = tm.base + (long long)(row + nbox * tm.box_rows + k) * tm.pitch + col;
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(d0 + (unsigned)(rim_elem + k * tm.box_cols) * 8u), "l"(src), "r"(rim_bytes), "r"(b)
                         : "memory");
        }
    }
    // All threads: wait until the copy announced for this phase has landed.  Bounded (a lost copy must not hang
    // the GPU); returns false on timeout so that the caller can raise the plan's error word.
    __device__ __forceinline__ bool mbar_wait(unsigned long long* bar, unsigned phase) const {
        const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
        unsigned ok = 0;
        for (int spins = 0; !ok && spins < (1 << 26); ++spins) {
            asm volatile(
                "{\n\t.reg .pred p;\n\t"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                "selp.u32 %0, 1, 0, p;\n\t}"
                : "=r"(ok)
                : "r"(a), "r"(phase)
                : "memory");
        }
        return ok != 0;
    }
};
#endif

}  // namespace litho
