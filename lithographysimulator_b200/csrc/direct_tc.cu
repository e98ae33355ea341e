// Direct ("Abbe") solver on the 5th-generation tensor cores -- reference imageformation.py:3-30.
//
//     E_s = A * G_s * A^T,  A[a][c] = w[c] exp(-i (2 pi / lambda) fp16(fp16(k[a]) fp16(x[c]))),  G_s = roll(P, s) . M
//
// is two complex matrix products per source point.  Both have the form
//     C[m][n] = sum_k A[m][kcol(k)] * Bt[n][k]          (A: the pn x pn operator, a window of its columns, wrapping)
//   stage 1:  U_s[b][u] = sum_v A[b][col(v)] G_s[u][v]   (Bt = G_s, built on the fly from pupil and mask spectrum)
//   stage 2:  E_s[a][b] = sum_u A[a][row(u)] U_s[b][u]   (Bt = U_s), epilogue |E_s|^2
// and run as REAL GEMMs on tcgen05.mma kind::tf32 with fp32 accumulation in tensor memory:
//   * a complex row of A is the K-vector (re, im, re, im, ...) as it lies in memory; every complex row of Bt
//     becomes two rows of the B operand, (re, -im, ...) and (im, re, ...), so that D[m][2n] = Re C, D[m][2n+1] = Im C
//     land in adjacent TMEM columns of the same lane (one thread owns a whole pixel: |E|^2 needs no exchange);
//   * 3xTF32: x = hi + lo with hi = rna_tf32(x); D += A_hi B_hi + A_lo B_hi + A_hi B_lo.  One TF32 pass misses
//     the 1e-5 parity bar by a factor of 20 (measured 2.2e-4, profiles/r03a_direct_bench.json), three passes keep
//     it at 1e-6 -- the dropped lo*lo term and the TF32 truncation of lo are O(2^-21).
// One CTA = one 128-row tile x one tile of <= 128 complex columns x one source point.  All 256 threads split the
// fp32 operands and write them to shared memory in the canonical K-major SWIZZLE_128B layout (32 tf32 = 128 bytes
// per row and k-block), one thread issues the 12 MMAs of a k-block, tcgen05.commit releases the stage through an
// mbarrier (2 stages: the fill of k-block j+1 overlaps the MMAs of k-block j), and the epilogue reads the
// accumulator with tcgen05.ld (32 lanes x 32 columns per warp and instruction).
#include <cuda_runtime.h>
#include <stdint.h>

#include "direct_tc.h"

namespace litho_tc {

constexpr int BM = 128;                 // rows per CTA = TMEM lanes
constexpr int KB = 16;                  // complex K elements per k-block (32 tf32 = one 128-byte swizzle row)
constexpr int THREADS = 256;            // two threads per operand row (4 of its 8 chunks each); 8 warps in the epilogue
constexpr int A_TILE = BM * 128;        // bytes of one A operand tile (hi or lo)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// K-major, SWIZZLE_128B, 8-row groups 1024 bytes apart (cute::UMMA::SmemDescriptor, version 1 = sm_100)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);     // start address, 16-byte units          bits [0,14)
    d |= (uint64_t)1 << 16;                     // leading byte offset (unused here)     bits [16,30)
    d |= (uint64_t)(1024 >> 4) << 32;           // stride byte offset between row groups bits [32,46)
    d |= (uint64_t)1 << 46;                     // descriptor version                    bits [46,48)
    d |= (uint64_t)2 << 61;                     // layout type SWIZZLE_128B              bits [61,64)
    return d;
}

__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}\n"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum), "r"(0u), "r"(0u), "r"(0u), "r"(0u)
        : "memory");
}

__device__ __forceinline__ uint32_t tf32_hi(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}

// element (row, 16-byte chunk c16) of a K-major SWIZZLE_128B tile whose base is 1024-byte aligned
__device__ __forceinline__ uint32_t swz(int row, int c16) { return (uint32_t)(row * 128 + ((c16 ^ (row & 7)) << 4)); }

__device__ __forceinline__ void st_pair(unsigned char* hi_tile, unsigned char* lo_tile, int row, int c16, float4 v) {
    const uint32_t h0 = tf32_hi(v.x), h1 = tf32_hi(v.y), h2 = tf32_hi(v.z), h3 = tf32_hi(v.w);
    const uint32_t o = swz(row, c16);
    *reinterpret_cast<uint4*>(hi_tile + o) = make_uint4(h0, h1, h2, h3);
    // the MMA truncates its 32-bit operands to TF32: round lo to nearest here, so that what is dropped is unbiased
    *reinterpret_cast<uint4*>(lo_tile + o) = make_uint4(tf32_hi(v.x - __uint_as_float(h0)), tf32_hi(v.y - __uint_as_float(h1)),
                                                        tf32_hi(v.z - __uint_as_float(h2)), tf32_hi(v.w - __uint_as_float(h3)));
}

__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    for (int spins = 0; !ok && spins < (1 << 24); ++spins) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    }
    return ok != 0;
}

__device__ __forceinline__ int imod(int a, int b) {
    int m = a % b;
    return m < 0 ? m + b : m;
}

// grid = (M tiles, N tiles, source points of the batch), block = 256
__global__ void __launch_bounds__(THREADS, 1) direct_tc_kernel(const __grid_constant__ TcParams P) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const int tid = threadIdx.x, warp = tid >> 5;
    const int row = tid & (BM - 1), half = tid >> 7;     // operand row of this thread, and which half of its chunks
    const int mt = blockIdx.x, nt = blockIdx.y, sl = blockIdx.z;
    const int pn = P.pn;
    const int NT = P.NT;                        // complex columns per N tile (multiple of 8, <= 128)
    const int NR = 2 * NT;                      // rows of the B operand = accumulator columns
    const uint32_t b_tile = (uint32_t)NR * 128; // bytes of one B operand tile (hi or lo)
    const uint32_t stage_bytes = 2 * A_TILE + 2 * b_tile;
    unsigned char* ctl = smem + 2 * stage_bytes;                 // mbarriers + TMEM address
    uint64_t* bars = reinterpret_cast<uint64_t*>(ctl);           // [0], [1]: stage free
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(ctl + 32);

    const int2 sh = P.shifts ? P.shifts[P.s_begin + sl] : make_int2(0, 0);
    const int Kc = P.stage == 1 ? P.Sc : P.Sr;                   // complex K extent
    const int kbase = P.stage == 1 ? P.pc0 + sh.y : P.pr0 + sh.x;   // grid column of A for k = 0 (wraps)
    const int nk = (Kc + KB - 1) / KB;

    if (tid == 0) {
        for (int i = 0; i < 2; ++i)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bars + i)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmem_slot;

    // instruction descriptor (cute::UMMA::InstrDescriptor): D = F32, A = B = TF32, both K-major, N = NR, M = 128
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NR >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

    // ---- this thread's rows of the two operands ----
    const int m = mt * BM + row;                 // A row (an output row of this stage)
    const cplx2* Arow = P.A + (size_t)(m < pn ? m : 0) * pn;
    const int n = nt * NT + row;                 // Bt row (an output column of this stage), rows >= NT idle in the B fill
    const bool n_fill = row < NT;
    const bool n_live = n_fill && n < (P.stage == 1 ? P.Sr : pn);
    const cplx2* Urow = P.U + ((size_t)sl * pn + (n_live ? n : 0)) * P.Upitch;                    // stage 2
    const cplx2* prow = P.pupil + (size_t)(P.pr0 + (n_live ? n : 0)) * pn + P.pc0;                // stage 1
    const cplx2* mrow = P.mask + (size_t)imod(P.pr0 + sh.x + (n_live ? n : 0), pn) * pn;          // stage 1
    bool lost = false;

    for (int j = 0; j < nk; ++j) {
        const int s = j & 1;
        unsigned char* st = smem + (size_t)s * stage_bytes;
        unsigned char* a_hi = st;
        unsigned char* a_lo = st + A_TILE;
        unsigned char* b_hi = st + 2 * A_TILE;
        unsigned char* b_lo = b_hi + b_tile;
        if (j >= 2 && !mbar_wait(smem_u32(bars + s), (uint32_t)(((j >> 1) - 1) & 1))) lost = true;   // MMAs of k-block j-2 done
        const int k0 = j * KB;
        // A tile: 16 complex per row = 8 chunks of 2, four per thread
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
            const int c = 4 * half + cc;
            const int k = k0 + 2 * c;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (m < pn) {
                if (k < Kc) { const cplx2 z = __ldg(Arow + imod(kbase + k, pn)); v.x = z.x; v.y = z.y; }
                if (k + 1 < Kc) { const cplx2 z = __ldg(Arow + imod(kbase + k + 1, pn)); v.z = z.x; v.w = z.y; }
            }
            st_pair(a_hi, a_lo, row, c, v);
        }
        // B tile: complex row n -> operand rows 2*row (re, -im) and 2*row+1 (im, re)
        if (n_fill) {
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) {
                const int c = 4 * half + cc;
                cplx2 z[2];
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int k = k0 + 2 * c + e;
                    z[e] = make_float2(0.f, 0.f);
                    if (n_live && k < Kc) {
                        if (P.stage == 1) {
                            const cplx2 p = __ldg(prow + k), q = __ldg(mrow + imod(P.pc0 + sh.y + k, pn));
                            z[e] = make_float2(p.x * q.x - p.y * q.y, p.x * q.y + p.y * q.x);
                        } else {
                            z[e] = Urow[k];
                        }
                    }
                }
                st_pair(b_hi, b_lo, 2 * row, c, make_float4(z[0].x, -z[0].y, z[1].x, -z[1].y));
                st_pair(b_hi, b_lo, 2 * row + 1, c, make_float4(z[0].y, z[0].x, z[1].y, z[1].x));
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the MMA
        __syncthreads();
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t ah = smem_u32(a_hi), al = smem_u32(a_lo), bh = smem_u32(b_hi), bl = smem_u32(b_lo);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {     // 4 MMAs of K = 8 tf32 (32 bytes) per 128-byte row
                const uint32_t off = kk * 32;
                mma_tf32(tmem, make_desc(ah + off), make_desc(bh + off), idesc, (j | kk) ? 1u : 0u);
                mma_tf32(tmem, make_desc(al + off), make_desc(bh + off), idesc, 1u);
                mma_tf32(tmem, make_desc(ah + off), make_desc(bl + off), idesc, 1u);
            }
            // arrives on the stage's mbarrier when every MMA issued so far has completed
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bars + s)) : "memory");
        }
    }
    // the last commit covers all MMAs
    if (!mbar_wait(smem_u32(bars + ((nk - 1) & 1)), (uint32_t)(((nk - 1) >> 1) & 1))) lost = true;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (lost && P.err) *P.err = 1;

    // ---- epilogue: a warp can reach the TMEM lanes 32*(warp % 4) .. +31; thread = one output row; warps 0-3 take
    // the even 32-column chunks, warps 4-7 the odd ones ----
    const uint32_t taddr = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    for (int c0 = 32 * half; c0 < NR; c0 += 64) {       // 32 accumulator columns = 16 complex outputs
        uint32_t r[32];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
              "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
              "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
              "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
            : "r"(taddr + (uint32_t)c0));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (m < pn) {
            const int nloc = c0 >> 1;            // first complex column of this chunk inside the tile
            const int nbase = nt * NT + nloc;
            // the last chunk of a tile narrower than a multiple of 16 columns reads past NR: those columns belong
            // to the next tile (another CTA), never store them
            const int lim = (NT - nloc) < 16 ? (NT - nloc) : 16;
            if (P.stage == 1) {
                cplx2* dst = P.Uout + ((size_t)sl * pn + m) * P.Upitch + nbase;
#pragma unroll
                for (int e = 0; e < 16; ++e)
                    if (e < lim && nbase + e < P.Sr) dst[e] = make_float2(__uint_as_float(r[2 * e]), __uint_as_float(r[2 * e + 1]));
            } else {
                float* dst = P.part + ((size_t)sl * pn + m) * pn + nbase;
#pragma unroll
                for (int e = 0; e < 16; ++e) {
                    const float x = __uint_as_float(r[2 * e]), y = __uint_as_float(r[2 * e + 1]);
                    if (e < lim && nbase + e < pn) dst[e] = x * x + y * y;
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u) : "memory");
}

// intensity[i] += sum_sl w[sl] * part[sl][i], slices in order (deterministic)
__global__ void direct_tc_reduce_kernel(const float* __restrict__ part, const float* __restrict__ weights, int s_begin,
                                        int batch, size_t plane, float* __restrict__ intensity) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= plane) return;
    float acc = 0.f;
    for (int sl = 0; sl < batch; ++sl) acc += (weights ? weights[s_begin + sl] : 1.f) * part[(size_t)sl * plane + i];
    intensity[i] += acc;
}

size_t tc_smem_bytes(int NT) { return (size_t)2 * (2 * A_TILE + 2 * (size_t)(2 * NT) * 128) + 64; }

int tc_tile_cols(int n) {   // complex columns per N tile: <= 128, a multiple of 8, tiles as even as possible
    const int tiles = (n + 127) / 128;
    const int per = (n + tiles - 1) / tiles;
    return (per + 7) / 8 * 8;
}

int tc_ctas_per_point(int pn, int n_out) {
    const int nt = tc_tile_cols(n_out);
    return ((pn + BM - 1) / BM) * ((n_out + nt - 1) / nt);
}

int tc_launch(TcParams P, int n_out, int batch, cudaStream_t st) {
    P.NT = tc_tile_cols(n_out);
    const size_t smem = tc_smem_bytes(P.NT);
    cudaError_t e = cudaFuncSetAttribute(direct_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return (int)e;
    dim3 grid((P.pn + BM - 1) / BM, (n_out + P.NT - 1) / P.NT, batch);
    direct_tc_kernel<<<grid, THREADS, smem, st>>>(P);
    return (int)cudaGetLastError();
}

int tc_reduce(const float* part, const float* weights, int s_begin, int batch, int pn, float* intensity, cudaStream_t st) {
    const size_t plane = (size_t)pn * pn;
    direct_tc_reduce_kernel<<<(unsigned)((plane + 255) / 256), 256, 0, st>>>(part, weights, s_begin, batch, plane, intensity);
    return (int)cudaGetLastError();
}

}  // namespace litho_tc
