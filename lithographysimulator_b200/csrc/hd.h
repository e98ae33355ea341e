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
// device-scope counters used by the fused persistent kernel
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
};
#endif

}  // namespace litho
