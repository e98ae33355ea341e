// C ABI of the Abbe imaging hot path (see include/litho_b200.h for the contract).
// Device build: nvcc -gencode arch=compute_100a,code=sm_100a.  With -DLITHO_EMU the same
// host logic drives the CPU emulation of the kernels (tests only).
#include "../../include/litho_b200.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "launch.h"
#include "direct_kernels.h"
#include "builders.h"

#if defined(LITHO_EMU)
#include "emu_runtime.h"
#else
#include "direct_tc.h"
#endif

using namespace litho;

// --------------------------------------------------------------------------- error state
static thread_local std::string g_err;
static int fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}

// --------------------------------------------------------------------------- backend
#if defined(LITHO_EMU)
static int be_malloc(void** p, size_t n) { *p = malloc(n ? n : 1); return *p ? 0 : 1; }
static void be_free(void* p) { free(p); }
static int be_h2d(void* d, const void* h, size_t n, litho_stream_t) { memcpy(d, h, n); return 0; }
static int be_d2h_sync(void* h, const void* d, size_t n, litho_stream_t) { memcpy(h, d, n); return 0; }
static const char* be_errstr(int) { return "emu"; }
#else
static int be_malloc(void** p, size_t n) { return (int)cudaMalloc(p, n ? n : 1); }
static void be_free(void* p) { cudaFree(p); }
static int be_h2d(void* d, const void* h, size_t n, litho_stream_t st) {
    return (int)cudaMemcpyAsync(d, h, n, cudaMemcpyHostToDevice, st);
}
static int be_d2h_sync(void* h, const void* d, size_t n, litho_stream_t st) {
    cudaError_t e = cudaMemcpyAsync(h, d, n, cudaMemcpyDeviceToHost, st);
    if (e != cudaSuccess) return (int)e;
    return (int)cudaStreamSynchronize(st);
}
static const char* be_errstr(int e) { return cudaGetErrorString((cudaError_t)e); }
#endif

#define BE_CHECK(expr)                                                                          \
    do {                                                                                        \
        int _e = (expr);                                                                        \
        if (_e != 0) return fail(LITHO_ERR_CUDA, std::string(#expr) + ": " + be_errstr(_e));    \
    } while (0)

// --------------------------------------------------------------------------- small kernels
namespace litho {

struct BBoxParams {
    const cplx* pupil;
    int pn;
    int* box;  // {rmin, rmax, cmin, cmax}
};

LITHO_HD void atomic_min_i(int* p, int v) {
#if defined(__CUDA_ARCH__)
    atomicMin(p, v);
#else
    if (v < *p) *p = v;
#endif
}
LITHO_HD void atomic_max_i(int* p, int v) {
#if defined(__CUDA_ARCH__)
    atomicMax(p, v);
#else
    if (v > *p) *p = v;
#endif
}

// one CTA per row; threads stride over the columns
template <class Ctx>
LITHO_HD void bbox_body(const BBoxParams& P, const Ctx& ctx) {
    const int row = ctx.bx();
    int cmin = P.pn, cmax = -1;
    for (int c = ctx.tid(); c < P.pn; c += ctx.bdim()) {
        cplx v = P.pupil[(size_t)row * P.pn + c];
        if (v.x != 0.f || v.y != 0.f) {
            if (c < cmin) cmin = c;
            if (c > cmax) cmax = c;
        }
    }
    if (cmax >= 0) {
        atomic_min_i(P.box + 0, row);
        atomic_max_i(P.box + 1, row);
        atomic_min_i(P.box + 2, cmin);
        atomic_max_i(P.box + 3, cmax);
    }
}

struct UnpermParams {
    PermView in;
    int pn;
    float* out;
};

// non-zero extents of the outermost `lines` bbox rows and columns of the pupil (absolute grid indices):
// ext[8k + ..] = {row r0+k: cmin,cmax | row r1-k: cmin,cmax | col c0+k: rmin,rmax | col c1-k: rmin,rmax}
struct ExtParams {
    const cplx* pupil;
    int pn, r0, r1, c0, c1, lines;
    int* ext;
};

template <class Ctx>
LITHO_HD void ext_body(const ExtParams& P, const Ctx& ctx) {
    for (int k = 0; k < P.lines; ++k) {
        int* e = P.ext + 8 * k;
        const int ra = P.r0 + k, rb = P.r1 - k, ca = P.c0 + k, cb = P.c1 - k;
        // rows and columns run out independently (a 1 x 35 window still has three column lines per side)
        const bool rows = ra >= 0 && rb < P.pn && (k == 0 || ra <= P.r1) && (k == 0 || rb >= P.r0);
        const bool cols = ca >= 0 && cb < P.pn && (k == 0 || ca <= P.c1) && (k == 0 || cb >= P.c0);
        for (int i = ctx.tid(); i < P.pn; i += ctx.bdim()) {
            if (rows) {
                const cplx a = P.pupil[(size_t)ra * P.pn + i], b = P.pupil[(size_t)rb * P.pn + i];
                if (a.x != 0.f || a.y != 0.f) { atomic_min_i(e + 0, i); atomic_max_i(e + 1, i); }
                if (b.x != 0.f || b.y != 0.f) { atomic_min_i(e + 2, i); atomic_max_i(e + 3, i); }
            }
            if (cols) {
                const cplx c = P.pupil[(size_t)i * P.pn + ca], d = P.pupil[(size_t)i * P.pn + cb];
                if (c.x != 0.f || c.y != 0.f) { atomic_min_i(e + 4, i); atomic_max_i(e + 5, i); }
                if (d.x != 0.f || d.y != 0.f) { atomic_min_i(e + 6, i); atomic_max_i(e + 7, i); }
            }
        }
    }
}

struct ShiftBoundsParams {
    const int2_* shifts;
    int n;
    int* out;  // {min d0, max d0, min d1, max d1}
};

template <class Ctx>
LITHO_HD void shift_bounds_body(const ShiftBoundsParams& P, const Ctx& ctx) {
    int lo0 = INT32_MAX, hi0 = INT32_MIN, lo1 = INT32_MAX, hi1 = INT32_MIN;
    for (int i = ctx.bx() * ctx.bdim() + ctx.tid(); i < P.n; i += ctx.bdim() * ctx.gdx()) {
        const int2_ s = P.shifts[i];
        lo0 = s.x < lo0 ? s.x : lo0; hi0 = s.x > hi0 ? s.x : hi0;
        lo1 = s.y < lo1 ? s.y : lo1; hi1 = s.y > hi1 ? s.y : hi1;
    }
    if (hi0 >= lo0) {
        atomic_min_i(P.out + 0, lo0); atomic_max_i(P.out + 1, hi0);
        atomic_min_i(P.out + 2, lo1); atomic_max_i(P.out + 3, hi1);
    }
}

// coarse plane [2][2][M][M] -> natural Nc x Nc plane (o = 2k + r on both axes)
struct CoarseUnpermParams {
    const float* ic;
    int M;
    float* out;  // [2M][2M]
};
LITHO_HD void coarse_unperm_elem(const CoarseUnpermParams& P, int a, int b) {
    const int M = P.M;
    P.out[(size_t)a * (2 * M) + b] = P.ic[((size_t)((a & 1) * 2 + (b & 1)) * M + (a >> 1)) * M + (b >> 1)];
}

#if !defined(LITHO_EMU)
__global__ void __launch_bounds__(256) fma_probe_kernel(float* out, int iters) {
    float a[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = (float)(threadIdx.x + i) * 1e-3f;
    const float b = 0.999f, c = 1e-4f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], b, c);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += a[i];
    if (s == 12345.678f) out[0] = s;  // never true: keeps the loop alive without memory traffic
}
// Source-point extraction in two short launches (imageformation.py:59: argwhere(lightsource) - pn//2, row-major
// order), for callers that stage one image after another next to persistent compute kernels, where every extra
// small kernel of the torch op sequence (nonzero, sub, cast, slice, bounds: ~10 launches) waits ~0.1 ms for a free
// SM slot.  The plane is cut into SP_CTAS contiguous chunks: pass 1 counts the source points of each chunk, pass 2
// gives every CTA the sum of the counts before it and compacts its chunk in order (warp ballots + an 8-entry shared
// prefix per 256-element sub-chunk; sub-chunks without a source point -- almost all -- cost one barrier).  Point
// number o (in reference order) goes to this rank iff o % world == rank, at index o / world.
#define SP_CTAS 296
struct SourcePointsParams {
    const unsigned char* ls;
    int elem_size, is_float, pn, rank, world, capacity;
    int2_* out;
    int* meta;     // {n_all, n_mine, min d0, max d0, min d1, max d1} (bounds over ALL points)
    int* counts;   // [SP_CTAS]
    size_t chunk;  // elements per CTA (a multiple of 256)
};
__device__ __forceinline__ bool sp_nonzero(const unsigned char* p, size_t i, int es, int is_float) {
    switch (es) {
        case 8: { const unsigned long long v = reinterpret_cast<const unsigned long long*>(p)[i];
                  return is_float ? (v << 1) != 0ull : v != 0ull; }
        case 4: { const unsigned v = reinterpret_cast<const unsigned*>(p)[i]; return is_float ? (v << 1) != 0u : v != 0u; }
        case 2: { const unsigned short v = reinterpret_cast<const unsigned short*>(p)[i];
                  return is_float ? (unsigned short)(v << 1) != 0 : v != 0; }
        default: return p[i] != 0;
    }
}
__global__ void __launch_bounds__(256) source_count_kernel(const __grid_constant__ SourcePointsParams P) {
    const size_t total = (size_t)P.pn * P.pn;
    const size_t lo = (size_t)blockIdx.x * P.chunk, hi = lo + P.chunk < total ? lo + P.chunk : total;
    int c = 0;
    for (size_t i = lo + threadIdx.x; i < hi; i += 256) c += sp_nonzero(P.ls, i, P.elem_size, P.is_float) ? 1 : 0;
    c = __reduce_add_sync(0xffffffffu, c);
    __shared__ int wc[8];
    if ((threadIdx.x & 31) == 0) wc[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        int s = 0;
        for (int j = 0; j < 8; ++j) s += wc[j];
        P.counts[blockIdx.x] = s;
    }
}
__global__ void __launch_bounds__(256) source_compact_kernel(const __grid_constant__ SourcePointsParams P) {
    __shared__ int warp_cnt[8];
    __shared__ int s_running;
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    const size_t total = (size_t)P.pn * P.pn;
    const size_t lo = (size_t)blockIdx.x * P.chunk, hi = lo + P.chunk < total ? lo + P.chunk : total;
    const int half = P.pn / 2;
    {   // points before this chunk, and (CTA 0) the grand total
        int before = 0, all = 0;
        for (int j = t; j < SP_CTAS; j += 256) {
            const int c = P.counts[j];
            all += c;
            before += j < (int)blockIdx.x ? c : 0;
        }
        before = __reduce_add_sync(0xffffffffu, before);
        all = __reduce_add_sync(0xffffffffu, all);
        __shared__ int wb[8], wa[8];
        if (lane == 0) { wb[w] = before; wa[w] = all; }
        __syncthreads();
        if (t == 0) {
            int sb = 0, sa = 0;
            for (int j = 0; j < 8; ++j) { sb += wb[j]; sa += wa[j]; }
            s_running = sb;
            if (blockIdx.x == 0) {
                P.meta[0] = sa;
                P.meta[1] = sa > P.rank ? (sa - P.rank + P.world - 1) / P.world : 0;
            }
        }
        __syncthreads();
    }
    if (P.counts[blockIdx.x] == 0) return;
    int lo0 = INT32_MAX, hi0 = INT32_MIN, lo1 = INT32_MAX, hi1 = INT32_MIN;
    for (size_t base = lo; base < hi; base += 256) {
        const size_t i = base + t;
        const bool f = i < hi && sp_nonzero(P.ls, i, P.elem_size, P.is_float);
        if (!__syncthreads_or((int)f)) continue;
        const unsigned b = __ballot_sync(0xffffffffu, f);
        if (lane == 0) warp_cnt[w] = __popc(b);
        __syncthreads();
        int before = 0, all = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = warp_cnt[j];
            before += j < w ? c : 0;
            all += c;
        }
        const int running = s_running;
        if (f) {
            const int o = running + before + __popc(b & ((1u << lane) - 1u));
            const int d0 = (int)(i / P.pn) - half, d1 = (int)(i % P.pn) - half;
            lo0 = min(lo0, d0); hi0 = max(hi0, d0); lo1 = min(lo1, d1); hi1 = max(hi1, d1);
            if (o % P.world == P.rank) {
                const int idx = o / P.world;
                if (idx < P.capacity) { P.out[idx].x = d0; P.out[idx].y = d1; }
            }
        }
        __syncthreads();
        if (t == 0) s_running = running + all;
        __syncthreads();
    }
    if (hi0 >= lo0) {
        atomicMin(P.meta + 2, lo0); atomicMax(P.meta + 3, hi0);
        atomicMin(P.meta + 4, lo1); atomicMax(P.meta + 5, hi1);
    }
}
__global__ void bbox_kernel(const __grid_constant__ BBoxParams P) { bbox_body(P, DevCtx{}); }
__global__ void ext_kernel(const __grid_constant__ ExtParams P) { ext_body(P, DevCtx{}); }
__global__ void shift_bounds_kernel(const __grid_constant__ ShiftBoundsParams P) { shift_bounds_body(P, DevCtx{}); }
__global__ void rim_kernel(const __grid_constant__ RimParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    rim_body(P, DevCtx{}, reinterpret_cast<cplx*>(smem_raw));
}
__global__ void rim_reduce_kernel(const __grid_constant__ RimReduceParams P) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < P.n) rim_reduce_elem(P, j);
}
__global__ void coarse_unperm_kernel(const __grid_constant__ CoarseUnpermParams P) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < 2 * P.M) coarse_unperm_elem(P, blockIdx.y, b);
}
__global__ void assemble_kernel(const __grid_constant__ AssembleParams P, int cols) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < cols) assemble_elem(P, blockIdx.y, j);
}
__global__ void finalize_kernel(const __grid_constant__ FinalizeParams P) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    if (x < P.out_side) finalize_pixel(P, y, x);
}
__global__ void direct_op_kernel(const __grid_constant__ DirectOpParams P) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < P.pn) direct_op_elem(P, blockIdx.y, c);
}
template <int KIND>
__global__ void __launch_bounds__(DIRECT_THREADS) direct_rows_kernel(const __grid_constant__ DirectParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    direct_rows_body<KIND>(P, DevCtx{}, reinterpret_cast<cplx*>(smem_raw));
}
template <int EPI>
__global__ void __launch_bounds__(DIRECT_THREADS) direct_cols_kernel(const __grid_constant__ DirectParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    direct_cols_body<EPI>(P, DevCtx{}, reinterpret_cast<cplx*>(smem_raw));
}
__global__ void source_kernel(const __grid_constant__ SourceParams P) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < P.pn) source_pixel(P, blockIdx.y, j);
}
__global__ void pupil_kernel(const __grid_constant__ PupilParams P) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < P.pn) pupil_pixel(P, blockIdx.y, j);
}
__global__ void resample_kernel(const __grid_constant__ ResampleParams P) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    if (x < P.side) resample_pixel(P, y, x);
}
__global__ void unpermute_kernel(const __grid_constant__ UnpermParams P) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    if (x < P.pn) P.out[(size_t)y * P.pn + x] = P.in.at(y, x);
}
#endif

}  // namespace litho

// --------------------------------------------------------------------------- twiddle tables
// w_L[i] = exp(+2*pi*i*i/L), computed in double on the host, one table per (device, L), kept for
// the life of the process (at most 128 KB each).
#include <map>
#include <mutex>
static std::mutex g_tw_mutex;
static std::map<std::pair<int, int>, cplx*> g_tw_cache;

static int get_twiddles(int L, cplx** out) {
    int dev = 0;
#if !defined(LITHO_EMU)
    if (cudaGetDevice(&dev) != cudaSuccess) return 1;
#endif
    std::lock_guard<std::mutex> lock(g_tw_mutex);
    auto key = std::make_pair(dev, L);
    auto it = g_tw_cache.find(key);
    if (it != g_tw_cache.end()) {
        *out = it->second;
        return 0;
    }
    std::vector<cplx> tw(L);
    for (int i = 0; i < L; ++i) {
        const double a = 2.0 * M_PI * (double)i / (double)L;
        tw[i].x = (float)cos(a);
        tw[i].y = (float)sin(a);
    }
    cplx* d = nullptr;
    int rc = be_malloc((void**)&d, sizeof(cplx) * L);
    if (rc == 0) rc = be_h2d(d, tw.data(), sizeof(cplx) * L, 0);
#if !defined(LITHO_EMU)
    if (rc == 0) rc = (int)cudaStreamSynchronize(0);
#endif
    if (rc != 0) {
        if (d) be_free(d);
        return rc;
    }
    g_tw_cache[key] = d;
    *out = d;
    return 0;
}

// --------------------------------------------------------------------------- small persistent scratch
// The support-analysis entry points (pupil_bbox, pupil_support, shift_bounds, pupil_build) need a few device
// words each and synchronise anyway; a per-device scratch that lives for the process replaces a cudaMalloc +
// cudaFree pair per call (cudaFree synchronises the whole device).  The mutex is held for the duration of the call.
static std::mutex g_scratch_mutex;
static std::map<int, int*> g_scratch;
static const size_t SCRATCH_BYTES = 16384;
static int get_scratch(int** out) {   // call with g_scratch_mutex held
    int dev = 0;
#if !defined(LITHO_EMU)
    if (cudaGetDevice(&dev) != cudaSuccess) return 1;
#endif
    auto it = g_scratch.find(dev);
    if (it != g_scratch.end()) {
        *out = it->second;
        return 0;
    }
    int* d = nullptr;
    const int rc = be_malloc((void**)&d, SCRATCH_BYTES);
    if (rc != 0) return rc;
    g_scratch[dev] = d;
    *out = d;
    return 0;
}

// --------------------------------------------------------------------------- TMA tile map of the T ring
// T (any number of [Sr][M] complex64 planes back to back) as a 2-D float32 tensor: 2*M floats per row.
// The CUtensorMap is encoded through the driver entry point obtained from the runtime (no libcuda link).
#if !defined(LITHO_EMU)
#include <cuda.h>
typedef CUresult (*encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static encode_tiled_fn get_encode_tiled() {
    static encode_tiled_fn fn = []() -> encode_tiled_fn {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            return nullptr;
        return (encode_tiled_fn)ptr;
    }();
    return fn;
}
#endif

// returns 0 on success; on failure the caller falls back to the plain-load column kernel
static int make_tile_map(TileMap* tm, const void* base, int M, long long rows, int box_rows, int box_cols) {
    memset(tm, 0, sizeof(*tm));
    tm->base = (const cplx*)base; tm->pitch = M; tm->rows = rows;
    tm->box_rows = box_rows; tm->box_cols = box_cols;
#if !defined(LITHO_EMU)
    static_assert(sizeof(CUtensorMap) == sizeof(tm->map), "CUtensorMap is 128 bytes");
    encode_tiled_fn enc = get_encode_tiled();
    if (!enc) return 1;
    if (((uintptr_t)base & 15) != 0 || rows <= 0 || rows > 0x7fffffffLL) return 1;
    const cuuint64_t gdim[2] = {(cuuint64_t)2 * M, (cuuint64_t)rows};
    const cuuint64_t gstr[1] = {(cuuint64_t)M * sizeof(cplx)};
    const cuuint32_t box[2] = {(cuuint32_t)2 * box_cols, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    CUtensorMapL2promotion l2 = CU_TENSOR_MAP_L2_PROMOTION_L2_128B;
    if (const char* env = getenv("LITHO_TMA_L2")) {
        const int v = atoi(env);
        l2 = v == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE : v == 64 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B
           : v == 256 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_L2_128B;
    }
    const CUresult r = enc((CUtensorMap*)tm->map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), gdim, gstr,
                           box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, l2,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return 1;
#endif
    return 0;
}

// --------------------------------------------------------------------------- plan
// T ring slots of the fast path: rows(b) may run LITHO_TSLOTS-1 batches ahead of cols(b)
#define LITHO_TSLOTS 3
// CTAs (= private slices) of the rim-sum kernel: each takes a contiguous chunk of the source points in order
#define LITHO_RIM_CHUNKS 128
// TMA-staged column pass: narrow tiles (two 256-thread CTAs per SM) or wide ones (one 512-thread CTA)
#ifndef LITHO_DEFAULT_COL_NARROW
#define LITHO_DEFAULT_COL_NARROW 1
#endif
struct litho_plan {
    int pn, N;
    int bbox[4];
    int Sr, Sc;
    ZoomPlan zp;
    AxisOut out;
    cplx* twL;  // device: w_L[i] = exp(+2*pi*i*i/L)
    int rows_fpc, cols_cb;
    int default_batch;
    // fast path (fast_kernels.h): coarse grid Nc = 2*Mf, q = N/Nc
    int path;        // 1 = generic fine grid, 2 = fast coarse grid
    int Mf, Nc, q, ppt;
    int er, ec;      // Sr-1-Mf / Sc-1-Mf: frequency lines Mf .. Mf+er (ec) need the rim sums; -1: none
    int ext[4][RIM_LINES][2];  // non-zero extents of the outermost window rows/columns (RimParams::ext)
    cplx* tables;    // device twiddle tables of the fast kernels (owned by the plan)
    int n_sm;
    int tma_cols;    // > 0: columns per tile of the TMA-staged column kernel (0: plain global loads)
    int tma_box_rows;  // rows per TMA box (TmaShape::BOX_ROWS)
    cplx* tables_c;  // device: compact twiddle tables of the TMA-staged column kernel (owned by the plan)
    cplx* tables_r;  // device: [tw1][tw1_odd] of the folded row pass (Mf = 1024; owned by the plan)
    // T ring bookkeeping across accumulate calls (LITHO_PHASE_INPUTS_READY): which ring the ev_cols events of
    // the last call refer to
    mutable const void* last_ws;
    mutable int last_batch;
    mutable int cols_recorded[LITHO_TSLOTS];
    int* status;     // device, 4 ints: [0] a shift had to be clamped, [1] a TMA tile copy never completed
    float* rim_scratch;   // device: LITHO_RIM_CHUNKS private slices of the rim sums (deterministic two-stage sum)
    size_t rim_stride;    // floats per slice
#if !defined(LITHO_EMU)
    cudaStream_t aux_stream;  // row passes of the fast path run here, overlapping the column passes
    cudaEvent_t ev_start, ev_rows[LITHO_TSLOTS], ev_cols[LITHO_TSLOTS];
#endif
};

static bool is_pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }

static int dispatch_fast_rows(int M, int ppt, const FastRowsParams& P, int gx, litho_stream_t st) {
    switch (M) {
#define X(m) case m: return launch_fast_rows_m<m, 32>(P, gx, st);
        LITHO_FOR_EACH_FAST_M(X)
#undef X
    }
    return -1;
}
static int dispatch_fast_cols(int M, int ppt, const FastColsParams& P, litho_stream_t st) {
    switch (M) {
#define X(m) case m: return launch_fast_cols_m<m, 32>(P, st);
        LITHO_FOR_EACH_FAST_M(X)
#undef X
    }
    return -1;
}
static int dispatch_fast_tma_cols(int M, int ppt, int which) {
    switch (M) {
#define X(m) case m: return fast_tma_cols_m<m, 32>(which);
        LITHO_FOR_EACH_FAST_M(X)
#undef X
    }
    return 0;
}
static int dispatch_fast_ntab(int M, int ppt) {
    switch (M) {
#define X(m) case m: return fast_ntab_m<m, 32>();
        LITHO_FOR_EACH_FAST_M(X)
#undef X
    }
    return -1;
}

// host mirror of FastShape<M,PPT>: table layout {pre[0..M], tw1[(t-1)*NS1+k], tw2[(t-1)*NS2+k]}
static std::vector<cplx> build_fast_tables(int M, int ppt) {
    int lg = 0;
    for (int m = M; m > 1; m >>= 1) ++lg;
    const int lp = ppt == 16 ? 4 : 5;
    const int R1 = lg > lp ? (lg - lp >= lp ? ppt : (1 << (lg - lp))) : 1;
    const int R2 = lg > 2 * lp ? (1 << (lg - 2 * lp)) : 1;
    const int NS1 = ppt, NS2 = ppt * R1;
    std::vector<cplx> t;
    auto push = [&](double num, double den) {
        const double a = 2.0 * M_PI * num / den;
        t.push_back(mk((float)cos(a), (float)sin(a)));
    };
    for (int u = 0; u <= M; ++u) push((double)u, 2.0 * M);
    for (int tt = 1; tt < R1; ++tt)
        for (int k = 0; k < NS1; ++k) push((double)tt * k, (double)NS1 * R1);
    for (int tt = 1; tt < R2; ++tt)
        for (int k = 0; k < NS2; ++k) push((double)tt * k, (double)NS2 * R2);
    return t;  // NTAB entries; the device copy is padded to an even count for 16-byte chunk copies
}

static int dispatch_shape(int M, int* fpc, int* cb) {
    switch (M) {
#define X(m) case m: shape_m<m>(fpc, cb); return 0;
        LITHO_FOR_EACH_M(X)
#undef X
    }
    return 1;
}
static int dispatch_rows(int M, int kind, const RowsParams& P, int gx, int gy, litho_stream_t st) {
    switch (M) {
#define X(m) case m: return launch_rows_m<m>(kind, P, gx, gy, st);
        LITHO_FOR_EACH_M(X)
#undef X
    }
    return -1;
}
static int dispatch_cols(int M, int epi, const ColsParams& P, int gx, int gy, litho_stream_t st) {
    switch (M) {
#define X(m) case m: return launch_cols_m<m>(epi, P, gx, gy, st);
        LITHO_FOR_EACH_M(X)
#undef X
    }
    return -1;
}

extern "C" {

int litho_abi_version(void) { return LITHO_ABI_VERSION; }
const char* litho_last_error(void) { return g_err.c_str(); }
int litho_is_device_build(void) {
#if defined(LITHO_EMU)
    return 0;
#else
    return 1;
#endif
}

int litho_epsilon_n(double deltaK, double pixelSize, double wavelength, double* eps, int* N) {
    if (!(deltaK > 0) || !(pixelSize > 0) || !(wavelength > 0)) return fail(LITHO_ERR_ARG, "epsilon_n: non-positive argument");
    const double beta = 1.0 / ((deltaK * pixelSize) / wavelength);
    // mask.py:63-65: argmin over {2,4,...,16384} of |2^k - beta| evaluated in float32, first wins
    int best = 2;
    float bestd = fabsf(2.0f - (float)beta);
    for (int k = 2; k <= 14; ++k) {
        const float d = fabsf((float)(1 << k) - (float)beta);
        if (d < bestd) {
            bestd = d;
            best = 1 << k;
        }
    }
    if (N) *N = best;
    if (eps) *eps = (double)best / beta;
    return LITHO_OK;
}

int litho_pupil_bbox(const void* pupil, int pn, int* bbox_host, void* stream) {
    if (!pupil || pn <= 0 || !bbox_host) return fail(LITHO_ERR_ARG, "pupil_bbox: bad argument");
    litho_stream_t st = (litho_stream_t)stream;
    std::lock_guard<std::mutex> scratch_lock(g_scratch_mutex);
    int* dbox = nullptr;
    BE_CHECK(get_scratch(&dbox));
    int init[4] = {pn, -1, pn, -1};
    int rc = be_h2d(dbox, init, sizeof(init), st);
    if (rc == 0) {
        BBoxParams P{(const cplx*)pupil, pn, dbox};
#if defined(LITHO_EMU)
        litho_emu::launch(pn, 1, 1, 32, 0, [&](const litho_emu::EmuCtx& c, char*) { bbox_body(P, c); });
#else
        bbox_kernel<<<pn, 256, 0, st>>>(P);
        rc = (int)cudaGetLastError();
#endif
    }
    if (rc == 0) rc = be_d2h_sync(bbox_host, dbox, 4 * sizeof(int), st);
    if (rc != 0) return fail(LITHO_ERR_CUDA, std::string("pupil_bbox: ") + be_errstr(rc));
    if (bbox_host[1] < 0) {
        bbox_host[0] = 0; bbox_host[1] = -1; bbox_host[2] = 0; bbox_host[3] = -1;
    }
    return LITHO_OK;
}

int litho_pupil_support(const void* pupil, int pn, int* support_host, void* stream) {
    return litho_pupil_support_lines(pupil, pn, 1, support_host, stream);
}

int litho_pupil_support_lines(const void* pupil, int pn, int lines, int* support_host, void* stream) {
    if (!support_host || lines < 1 || lines > RIM_LINES) return fail(LITHO_ERR_ARG, "pupil_support: bad argument");
    int rc = litho_pupil_bbox(pupil, pn, support_host, stream);
    if (rc) return rc;
    const int r0 = support_host[0], r1 = support_host[1], c0 = support_host[2], c1 = support_host[3];
    const int n = 8 * lines;
    for (int i = 0; i < n; ++i) support_host[4 + i] = (i & 1) ? -1 : pn;   // empty extents: lo = pn, hi = -1
    if (r1 < r0) {
        for (int i = 0; i < n; ++i) support_host[4 + i] = (i & 1) ? -1 : 0;
        return LITHO_OK;
    }
    litho_stream_t st = (litho_stream_t)stream;
    std::lock_guard<std::mutex> scratch_lock(g_scratch_mutex);
    int* dext = nullptr;
    BE_CHECK(get_scratch(&dext));
    rc = be_h2d(dext, support_host + 4, n * sizeof(int), st);
    if (rc == 0) {
        ExtParams P{(const cplx*)pupil, pn, r0, r1, c0, c1, lines, dext};
#if defined(LITHO_EMU)
        litho_emu::launch(1, 1, 1, 32, 0, [&](const litho_emu::EmuCtx& c, char*) { ext_body(P, c); });
#else
        ext_kernel<<<1, 1024, 0, st>>>(P);
        rc = (int)cudaGetLastError();
#endif
    }
    if (rc == 0) rc = be_d2h_sync(support_host + 4, dext, n * sizeof(int), st);
    if (rc != 0) return fail(LITHO_ERR_CUDA, std::string("pupil_support: ") + be_errstr(rc));
    return LITHO_OK;
}

int litho_shift_bounds(const int32_t* shifts, int n_src, int* bounds_host, void* stream) {
    if (!bounds_host || n_src < 0 || (n_src > 0 && !shifts)) return fail(LITHO_ERR_ARG, "shift_bounds: bad argument");
    bounds_host[0] = bounds_host[2] = 0;
    bounds_host[1] = bounds_host[3] = 0;
    if (n_src == 0) return LITHO_OK;
    litho_stream_t st = (litho_stream_t)stream;
    std::lock_guard<std::mutex> scratch_lock(g_scratch_mutex);
    int* d = nullptr;
    BE_CHECK(get_scratch(&d));
    int init[4] = {INT32_MAX, INT32_MIN, INT32_MAX, INT32_MIN};
    int rc = be_h2d(d, init, sizeof(init), st);
    if (rc == 0) {
        ShiftBoundsParams P{(const int2_*)shifts, n_src, d};
#if defined(LITHO_EMU)
        litho_emu::launch(1, 1, 1, 32, 0, [&](const litho_emu::EmuCtx& c, char*) { shift_bounds_body(P, c); });
#else
        shift_bounds_kernel<<<(n_src + 255) / 256 > 64 ? 64 : (n_src + 255) / 256, 256, 0, st>>>(P);
        rc = (int)cudaGetLastError();
#endif
    }
    if (rc == 0) rc = be_d2h_sync(bounds_host, d, 4 * sizeof(int), st);
    if (rc != 0) return fail(LITHO_ERR_CUDA, std::string("shift_bounds: ") + be_errstr(rc));
    return LITHO_OK;
}

int litho_source_points(const void* lightsource, int elem_size, int is_float, int pn, int rank, int world,
                        int32_t* shifts, int capacity, int* meta_host, void* stream) {
    if (!lightsource || !meta_host || pn < 1 || world < 1 || rank < 0 || rank >= world || capacity < 0 ||
        (capacity > 0 && !shifts) || (elem_size != 1 && elem_size != 2 && elem_size != 4 && elem_size != 8))
        return fail(LITHO_ERR_ARG, "source_points: bad argument");
    litho_stream_t st = (litho_stream_t)stream;
    int init[6] = {0, 0, INT32_MAX, INT32_MIN, INT32_MAX, INT32_MIN};
#if defined(LITHO_EMU)
    (void)st;
    const unsigned char* p = (const unsigned char*)lightsource;
    int n = 0, mine = 0;
    for (size_t i = 0; i < (size_t)pn * pn; ++i) {
        bool nz = false;
        if (is_float && elem_size == 8) nz = ((const double*)p)[i] != 0.0;
        else if (is_float && elem_size == 4) nz = ((const float*)p)[i] != 0.0f;
        else if (is_float && elem_size == 2) nz = (unsigned short)(((const unsigned short*)p)[i] << 1) != 0;
        else for (int b = 0; b < elem_size; ++b) nz = nz || p[i * elem_size + b] != 0;
        if (!nz) continue;
        const int d0 = (int)(i / pn) - pn / 2, d1 = (int)(i % pn) - pn / 2;
        if (d0 < init[2]) init[2] = d0;
        if (d0 > init[3]) init[3] = d0;
        if (d1 < init[4]) init[4] = d1;
        if (d1 > init[5]) init[5] = d1;
        if (n % world == rank) {
            if (mine < capacity) { shifts[2 * mine] = d0; shifts[2 * mine + 1] = d1; }
            ++mine;
        }
        ++n;
    }
    init[0] = n; init[1] = mine;
    memcpy(meta_host, init, sizeof(init));
#else
    std::lock_guard<std::mutex> scratch_lock(g_scratch_mutex);
    int* dmeta = nullptr;
    BE_CHECK(get_scratch(&dmeta));      // scratch: meta (6 ints, padded to 16), then the per-chunk counts
    BE_CHECK(be_h2d(dmeta, init, sizeof(init), st));
    SourcePointsParams P;
    memset(&P, 0, sizeof(P));
    P.ls = (const unsigned char*)lightsource; P.elem_size = elem_size; P.is_float = is_float; P.pn = pn;
    P.rank = rank; P.world = world; P.capacity = capacity; P.out = (int2_*)shifts; P.meta = dmeta;
    P.counts = dmeta + 16;
    const size_t total = (size_t)pn * pn;
    P.chunk = ((total + SP_CTAS - 1) / SP_CTAS + 255) / 256 * 256;
    source_count_kernel<<<SP_CTAS, 256, 0, st>>>(P);
    BE_CHECK((int)cudaGetLastError());
    source_compact_kernel<<<SP_CTAS, 256, 0, st>>>(P);
    BE_CHECK((int)cudaGetLastError());
    BE_CHECK(be_d2h_sync(meta_host, dmeta, sizeof(init), st));
#endif
    if (meta_host[0] == 0) meta_host[2] = meta_host[3] = meta_host[4] = meta_host[5] = 0;
    return LITHO_OK;
}

int litho_plan_create(int pn, int N, const int* bbox, int flags, litho_plan_t** out) {
    if (!bbox) return fail(LITHO_ERR_ARG, "plan_create: null argument");
    return litho_plan_create_lines(pn, N, bbox, 0, flags, out);
}

int litho_plan_create_ex(int pn, int N, const int* support, int flags, litho_plan_t** out) {
    return litho_plan_create_lines(pn, N, support, 1, flags, out);
}

// support = bbox (4 ints) followed by the extents of `lines` outermost rows/columns (8 ints per line, the layout
// litho_pupil_support_lines returns).  Lines without measured extents are assumed to span the whole window
// (correct, just slower).
int litho_plan_create_lines(int pn, int N, const int* bbox, int lines, int flags, litho_plan_t** out) {
    if (!out || !bbox) return fail(LITHO_ERR_ARG, "plan_create: null argument");
    if (lines < 0 || lines > RIM_LINES) return fail(LITHO_ERR_ARG, "plan_create: lines out of range");
    if (pn < 2 || (pn & 1)) return fail(LITHO_ERR_ARG, "plan_create: pixelNumber must be even and >= 2 (odd grids make the reference transform length N-1)");
    if (!is_pow2(N) || N > 16384 || N < 16) return fail(LITHO_ERR_ARG, "plan_create: N must be a power of two in [16,16384]");
    if (N < pn) return fail(LITHO_ERR_ARG, "plan_create: N < pixelNumber is unsupported (the reference raises, SURVEY Q7)");
    int r0 = bbox[0], r1 = bbox[1], c0 = bbox[2], c1 = bbox[3];
    if (r1 < r0 || c1 < c0) {  // empty pupil: keep a 1x1 window so the kernels have something to do
        r0 = r1 = c0 = c1 = pn / 2;
    }
    if (r0 < 0 || c0 < 0 || r1 >= pn || c1 >= pn) return fail(LITHO_ERR_ARG, "plan_create: bbox outside the grid");
    litho_plan* p = new litho_plan();
    p->pn = pn; p->N = N;
    p->bbox[0] = r0; p->bbox[1] = r1; p->bbox[2] = c0; p->bbox[3] = c1;
    p->Sr = r1 - r0 + 1;
    p->Sc = c1 - c0 + 1;
    const int S = p->Sr > p->Sc ? p->Sr : p->Sc;
    int M = 16;
    while (M < S - 1) M <<= 1;
    if (M > N) M = N;
    p->zp.L = N; p->zp.M = M; p->zp.R = N / M;
    p->out.W = pn; p->out.center = pn / 2;
    p->zp.Wr = (pn + p->zp.R - 1) / p->zp.R;
    if (dispatch_shape(M, &p->rows_fpc, &p->cols_cb)) {
        delete p;
        return fail(LITHO_ERR_ARG, "plan_create: unsupported sub-FFT length");
    }
    p->twL = nullptr;
    int rc = get_twiddles(N, &p->twL);
    if (rc != 0) {
        delete p;
        return fail(LITHO_ERR_CUDA, std::string("plan_create: twiddle table: ") + be_errstr(rc));
    }
    size_t per = (size_t)p->zp.R * p->Sr * p->zp.Wr * sizeof(cplx);
    // ---- fast path eligibility: even window fit S <= Mf+1, coarse grid Nc = 2*Mf no finer than N ----
    p->path = 1; p->tables = nullptr; p->tables_c = nullptr; p->tables_r = nullptr; p->tma_box_rows = 0; p->Mf = p->Nc = p->q = 0; p->er = p->ec = -1;
    p->tma_cols = 0;
    p->status = nullptr; p->rim_scratch = nullptr; p->rim_stride = 0;
    p->last_ws = nullptr; p->last_batch = 0;
    for (int i = 0; i < LITHO_TSLOTS; ++i) p->cols_recorded[i] = 0;
#if !defined(LITHO_EMU)
    p->aux_stream = nullptr;
#endif
    p->n_sm = 148;
#if !defined(LITHO_EMU)
    {
        int dev = 0, n = 0;
        if (cudaGetDevice(&dev) == cudaSuccess &&
            cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
            p->n_sm = n;
    }
#endif
    // fast path: window fits S <= Mf + 1 + RIM_EXTRA (the inputs beyond Mf fold onto the first slots and the
    // frequency lines they add are carried by the rim sums)
    int Mf = 32;
    while (Mf + RIM_EXTRA < S - 1) Mf <<= 1;
    if (!(flags & LITHO_PLAN_GENERIC) && Mf <= 4096 && 2 * Mf <= N) {
        // points per thread of the fast FFTs: 32 (one exchange, 128 regs) or 16 (two exchanges, 64 regs,
        // twice the resident warps); LITHO_FAST_PPT overrides the per-size default for experiments
        p->ppt = 32;
        std::vector<cplx> tab = build_fast_tables(Mf, p->ppt);
        if ((int)tab.size() != dispatch_fast_ntab(Mf, p->ppt)) {  // (checked before the padding entry is added)
            delete p;
            return fail(LITHO_ERR_ARG, "plan_create: internal table layout mismatch");
        }
        tab.push_back(mk(0.f, 0.f));  // padding (NTAB_PAD)
        rc = be_malloc((void**)&p->tables, tab.size() * sizeof(cplx));
        if (rc == 0) rc = be_h2d(p->tables, tab.data(), tab.size() * sizeof(cplx), 0);
#if !defined(LITHO_EMU)
        if (rc == 0) rc = (int)cudaStreamSynchronize(0);
#endif
        if (rc != 0) {
            if (p->tables) be_free(p->tables);
            delete p;
            return fail(LITHO_ERR_CUDA, std::string("plan_create: fast tables: ") + be_errstr(rc));
        }
#if !defined(LITHO_EMU)
        {
            cudaError_t e = cudaStreamCreateWithFlags(&p->aux_stream, cudaStreamNonBlocking);
            if (e == cudaSuccess) e = cudaEventCreateWithFlags(&p->ev_start, cudaEventDisableTiming);
            for (int i = 0; i < LITHO_TSLOTS && e == cudaSuccess; ++i) {
                e = cudaEventCreateWithFlags(&p->ev_rows[i], cudaEventDisableTiming);
                if (e == cudaSuccess) e = cudaEventCreateWithFlags(&p->ev_cols[i], cudaEventDisableTiming);
            }
            if (e != cudaSuccess) {
                be_free(p->tables);
                delete p;
                return fail(LITHO_ERR_CUDA, std::string("plan_create: stream/event: ") + cudaGetErrorString(e));
            }
        }
#endif
        p->path = 2; p->Mf = Mf; p->Nc = 2 * Mf; p->q = N / (2 * Mf);
        // TMA-staged column kernel where the shape has one (32 <= M <= 4096).  LITHO_TMA=0 selects plain loads,
        // LITHO_COL_NARROW=0/1 the wide (one 512-thread CTA per SM) or narrow (two 256-thread CTAs) tile.
        {
            // measured (profiles/README.md, r02f/r02g): narrow wins at M = 1024 (+3.8 %) and M = 128 (+9 %),
            // wide at M = 512 (+3.5 %)
            int narrow = LITHO_DEFAULT_COL_NARROW && Mf != 512;
            if (const char* env = getenv("LITHO_COL_NARROW")) narrow = atoi(env) != 0;
            const int wide_c = dispatch_fast_tma_cols(Mf, p->ppt, 0), narrow_c = dispatch_fast_tma_cols(Mf, p->ppt, 1);
            p->tma_cols = (narrow && narrow_c > 0) ? narrow_c : wide_c;
            p->tma_box_rows = dispatch_fast_tma_cols(Mf, p->ppt, 2);
            if (const char* env = getenv("LITHO_TMA")) {
                if (atoi(env) == 0) p->tma_cols = 0;
            }
        }
        if (p->tma_cols > 0 || Mf >= 4096) {   // (Mf >= 4096: the row pass uses the compact tables too)
            // compact tables: pre[0..M/2] (padded to an even count), then tw1 and tw2 as in the full layout
            std::vector<cplx> tc(tab.begin(), tab.begin() + Mf / 2 + 1);
            if (tc.size() & 1) tc.push_back(mk(0.f, 0.f));
            tc.insert(tc.end(), tab.begin() + Mf + 1, tab.end() - 1);  // (the last entry of tab is its padding)
            if (tc.size() & 1) tc.push_back(mk(0.f, 0.f));
            if ((int)tc.size() != dispatch_fast_tma_cols(Mf, p->ppt, 3)) {
                be_free(p->tables);
                delete p;
                return fail(LITHO_ERR_ARG, "plan_create: internal compact table layout mismatch");
            }
            rc = be_malloc((void**)&p->tables_c, tc.size() * sizeof(cplx));
            if (rc == 0) rc = be_h2d(p->tables_c, tc.data(), tc.size() * sizeof(cplx), 0);
#if !defined(LITHO_EMU)
            if (rc == 0) rc = (int)cudaStreamSynchronize(0);
#endif
            if (rc != 0) {
                be_free(p->tables);
                if (p->tables_c) be_free(p->tables_c);
    if (p->tables_r) be_free(p->tables_r);
                delete p;
                return fail(LITHO_ERR_CUDA, std::string("plan_create: compact tables: ") + be_errstr(rc));
            }
        }
        if (Mf == 1024 && p->ppt == 32) {
            // folded row pass (FastShape::ROW_FOLD): tw1[t][k] = w_M^(t k) and tw1_odd[t][k] = w_2M^(t (2k+1)), t = 1..31
            std::vector<cplx> tr;
            for (int odd = 0; odd < 2; ++odd)
                for (int t = 1; t < 32; ++t)
                    for (int k = 0; k < 32; ++k) {
                        const double a = 2.0 * M_PI * (double)(t * (2 * k + odd)) / (2.0 * Mf);
                        tr.push_back(mk((float)cos(a), (float)sin(a)));
                    }
            rc = be_malloc((void**)&p->tables_r, tr.size() * sizeof(cplx));
            if (rc == 0) rc = be_h2d(p->tables_r, tr.data(), tr.size() * sizeof(cplx), 0);
#if !defined(LITHO_EMU)
            if (rc == 0) rc = (int)cudaStreamSynchronize(0);
#endif
            if (rc != 0) {
                const std::string msg = std::string("plan_create: folded row tables: ") + be_errstr(rc);
                litho_plan_destroy(p);
                return fail(LITHO_ERR_CUDA, msg);
            }
        }
        p->er = p->Sr > Mf ? p->Sr - 1 - Mf : -1;
        p->ec = p->Sc > Mf ? p->Sc - 1 - Mf : -1;
        // extents in window coordinates, clamped to the window; full width where none were measured
        auto clampi = [](int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); };
        for (int k = 0; k < RIM_LINES; ++k) {
            for (int side = 0; side < 4; ++side) {
                const int len = side < 2 ? p->Sc : p->Sr, org = side < 2 ? c0 : r0;
                int lo = 0, hi = len - 1;
                if (k < lines) {
                    lo = bbox[4 + 8 * k + 2 * side] - org;
                    hi = bbox[4 + 8 * k + 2 * side + 1] - org;
                    if (hi >= lo) { lo = clampi(lo, 0, len - 1); hi = clampi(hi, 0, len - 1); }
                    else { lo = 0; hi = -1; }   // an empty line (possible for k > 0)
                }
                p->ext[side][k][0] = lo; p->ext[side][k][1] = hi;
            }
        }
        per = (size_t)2 * p->Sr * Mf * sizeof(cplx);
        // error words + the private slices of the two-stage rim sums (only when rim lines exist)
        rc = be_malloc((void**)&p->status, 4 * sizeof(int));
        if (rc == 0) {
            const int zero[4] = {0, 0, 0, 0};
            rc = be_h2d(p->status, zero, sizeof(zero), 0);
        }
        if (rc == 0 && p->q > 1 && (p->er >= 0 || p->ec >= 0)) {
            p->rim_stride = (size_t)2 * RIM_LINES * (2 * p->Sc - 1) + (size_t)2 * RIM_LINES * (2 * p->Sr - 1);
            rc = be_malloc((void**)&p->rim_scratch, (size_t)LITHO_RIM_CHUNKS * p->rim_stride * sizeof(float));
        }
#if !defined(LITHO_EMU)
        if (rc == 0) rc = (int)cudaStreamSynchronize(0);
#endif
        if (rc != 0) {
            const std::string msg = std::string("plan_create: status / rim scratch: ") + be_errstr(rc);
            litho_plan_destroy(p);
            return fail(LITHO_ERR_CUDA, msg);
        }
    }
    // batch (source points per launch pair).  Generic path: T of one batch around 64 MB.  Fast path:
    // measured on B200 at cfg3 (profiles/README.md): per-launch costs (table loads, the read-modify-write of
    // the intensity plane, partial waves, launch gaps) outweigh keeping T in L2 -- 58 / 67 / 73 / 76 / 77
    // images/s at batches of 4 / 8 / 16 / 32 / 48 -- so aim at ~800 MB per ring slot, at most 48 points.
    const size_t target = (p->path == 2) ? ((size_t)816 << 20) : ((size_t)64 << 20);
    int b = (int)(target / (per ? per : 1));
    const int cap = (p->path == 2) ? 48 : 16;
    if (p->path == 2) {
        // large grids: every column-pass launch re-reads and re-writes the intensity plane (268 MB at 8192 px), so a
        // launch must cover enough source points to amortise it -- measured at cfg5 (profiles/README.md, r03n):
        // 571 / 483 / 463 / 439 / 424 ms per image at batches of 3 / 6 / 8 / 12 / 16.  At least 16 points while the
        // 3-slot ring stays below 16 GB.
        int floor_b = 16;
        while (floor_b > 1 && (size_t)LITHO_TSLOTS * floor_b * per > ((size_t)16 << 30)) floor_b >>= 1;
        if (b < floor_b) b = floor_b;
    }
    p->default_batch = b < 1 ? 1 : (b > cap ? cap : b);
    *out = p;
    return LITHO_OK;
}

void litho_plan_destroy(litho_plan_t* p) {
    if (!p) return;
    if (p->tables) be_free(p->tables);
    if (p->tables_c) be_free(p->tables_c);
    if (p->status) be_free(p->status);
    if (p->rim_scratch) be_free(p->rim_scratch);
#if !defined(LITHO_EMU)
    if (p->aux_stream) {
        cudaStreamSynchronize(p->aux_stream);
        cudaStreamDestroy(p->aux_stream);
        cudaEventDestroy(p->ev_start);
        for (int i = 0; i < LITHO_TSLOTS; ++i) {
            cudaEventDestroy(p->ev_rows[i]);
            cudaEventDestroy(p->ev_cols[i]);
        }
    }
#endif
    delete p;  // the w_L twiddle table belongs to the process-wide cache
}

// coarse plane [2][2][Mf][Mf], then the rim sums: RIM_LINES row lines of 2*Sc-1 complex, RIM_LINES column lines
// of 2*Sr-1 complex (RimParams::frow / fcol)
static uint64_t rim_row_floats(const litho_plan* p) { return (uint64_t)2 * RIM_LINES * (2 * p->Sc - 1); }
static uint64_t rim_col_floats(const litho_plan* p) { return (uint64_t)2 * RIM_LINES * (2 * p->Sr - 1); }
static uint64_t intensity_elems(const litho_plan* p) {
    if (p->path == 2) return (uint64_t)4 * p->Mf * p->Mf + rim_row_floats(p) + rim_col_floats(p);
    return (uint64_t)p->zp.R * p->zp.R * p->zp.Wr * p->zp.Wr;
}

static float* rim_row_ptr(const litho_plan* p, float* intensity) { return intensity + (size_t)4 * p->Mf * p->Mf; }
static float* rim_col_ptr(const litho_plan* p, float* intensity) { return rim_row_ptr(p, intensity) + rim_row_floats(p); }

int litho_plan_get_info(const litho_plan_t* p, litho_plan_info_t* info) {
    if (!p || !info) return fail(LITHO_ERR_ARG, "plan_get_info: null argument");
    info->pn = p->pn; info->N = p->N;
    memcpy(info->bbox, p->bbox, sizeof(p->bbox));
    info->L = p->zp.L; info->M = p->zp.M; info->R = p->zp.R; info->Wr = p->zp.Wr;
    info->path = p->path;
    if (p->path == 2) {
        info->L = p->Nc; info->M = p->Mf; info->R = 2; info->Wr = p->Mf;
    }
    info->default_batch = p->default_batch;
    info->intensity_elems = intensity_elems(p);
    // shifts for which roll() does not wrap the pupil window (precondition of the fast path)
    info->shift_range[0] = -p->bbox[0]; info->shift_range[1] = p->pn - 1 - p->bbox[1];
    info->shift_range[2] = -p->bbox[2]; info->shift_range[3] = p->pn - 1 - p->bbox[3];
    return LITHO_OK;
}

int litho_plan_status(const litho_plan_t* p, int* status_host, void* stream) {
    if (!p || !status_host) return fail(LITHO_ERR_ARG, "plan_status: null argument");
    status_host[0] = status_host[1] = 0;
    if (!p->status) return LITHO_OK;   // generic plans have nothing to report
    int tmp[4] = {0, 0, 0, 0};
    BE_CHECK(be_d2h_sync(tmp, p->status, sizeof(tmp), (litho_stream_t)stream));
    status_host[0] = tmp[0]; status_host[1] = tmp[1];
    if (tmp[0] || tmp[1]) {
        const int zero[4] = {0, 0, 0, 0};
        BE_CHECK(be_h2d(p->status, zero, sizeof(zero), (litho_stream_t)stream));
#if !defined(LITHO_EMU)
        BE_CHECK((int)cudaStreamSynchronize((litho_stream_t)stream));
#endif
    }
    return LITHO_OK;
}

int litho_plan_column_tile(const litho_plan_t* p) {
    if (!p || p->path != 2 || p->tma_cols <= 0) return 0;
#if !defined(LITHO_EMU)
    if (!get_encode_tiled()) return 0;
#endif
    return p->tma_cols;
}

static size_t align16(size_t b) { return (b + 15) / 16 * 16; }

// T of the generic kernels (also used by litho_fft_field on any plan)
static size_t generic_t_bytes(const litho_plan* p, int batch) {
    return (size_t)batch * p->zp.R * p->Sr * p->zp.Wr * sizeof(cplx);
}

// default source points per launch pair when n_focus pupils are batched: the T ring holds n_focus blocks per point
static int default_batch_focus(const litho_plan* p, int n_focus) {
    if (n_focus <= 1) return p->default_batch;
    int b = p->default_batch / n_focus;
    return b < 1 ? 1 : b;
}

size_t litho_plan_workspace_bytes_focus(const litho_plan_t* p, int batch, int n_focus) {
    if (!p || n_focus < 1) return 0;
    if (batch <= 0) batch = default_batch_focus(p, n_focus);
    if (p->path == 2) {
        const size_t fast = (size_t)LITHO_TSLOTS * batch * n_focus * 2 * p->Sr * p->Mf * sizeof(cplx);  // T slots (ring)
        const size_t gen1 = generic_t_bytes(p, 1);
        return fast > gen1 ? fast : gen1;
    }
    return generic_t_bytes(p, batch);   // generic kernels: one focus value at a time
}

size_t litho_plan_workspace_bytes(const litho_plan_t* p, int batch) { return litho_plan_workspace_bytes_focus(p, batch, 1); }

// coarse -> fine interpolation buffers of the fast path (see fine_plane_fast)
struct FinalizeLayout {
    size_t a_off, t1_off, fhat_off, spec_off, t2_off, fine_off, total;
    ZoomPlan z1, z2;
    AxisOut o1, o2;
    int fpc1, cb1, fpc2, cb2;
    int spec_rows, spec_cols;
};

static int finalize_layout(const litho_plan* p, FinalizeLayout* L) {
    memset(L, 0, sizeof(*L));
    const int pn = p->pn;
    size_t off = 0;
    if (p->path == 2 && p->q > 1) {
        const int Nc = p->Nc, N = p->N;
        // step 1: centred forward DFT of the Nc x Nc coarse plane (one length-Nc FFT per line)
        L->z1.L = Nc; L->z1.M = Nc; L->z1.R = 1; L->z1.Wr = Nc;
        L->o1.W = Nc; L->o1.center = Nc / 2;
        // step 2: inverse zoom of the (Nc+1+2Er) x (Nc+1+2Ec) spectrum to the pn centre pixels of the N grid;
        // a spectrum wider than Nc+1 needs the next sub-FFT length (2*Nc <= N because q > 1)
        const int Er = p->er > 0 ? p->er : 0, Ec = p->ec > 0 ? p->ec : 0;
        const int M2 = (Er > 0 || Ec > 0) ? 2 * Nc : Nc;
        L->spec_rows = Nc + 1 + 2 * Er; L->spec_cols = Nc + 1 + 2 * Ec;
        L->z2.L = N; L->z2.M = M2; L->z2.R = N / M2; L->z2.Wr = (pn + L->z2.R - 1) / L->z2.R;
        L->o2.W = pn; L->o2.center = pn / 2;
        if (dispatch_shape(Nc, &L->fpc1, &L->cb1)) return 1;
        if (dispatch_shape(M2, &L->fpc2, &L->cb2)) return 1;
        L->a_off = off;    off += align16((size_t)Nc * Nc * sizeof(float));
        L->t1_off = off;   off += align16((size_t)Nc * Nc * sizeof(cplx));
        L->fhat_off = off; off += align16((size_t)Nc * Nc * sizeof(cplx));
        L->spec_off = off; off += align16((size_t)L->spec_rows * L->spec_cols * sizeof(cplx));
        L->t2_off = off;   off += align16((size_t)L->z2.R * L->spec_rows * L->z2.Wr * sizeof(cplx));
    }
    L->fine_off = off;
    off += align16((size_t)pn * pn * sizeof(float));
    L->total = off;
    return 0;
}

size_t litho_plan_finalize_workspace_bytes(const litho_plan_t* p) {
    if (!p) return 0;
    if (p->path != 2 || p->q == 1) return 16;  // nothing to stage
    FinalizeLayout L;
    if (finalize_layout(p, &L)) return 0;
    return L.total;
}

static AxisIn axis_in(int first, int pn, int S) {
    AxisIn a;
    a.first = first; a.period = pn; a.center = pn / 2; a.S = S;
    return a;
}

int litho_abbe_fft_accumulate(const litho_plan_t* p, const void* maskFT, const void* pupil, const int32_t* shifts,
                              const float* weights, int n_src, int batch, float* intensity, void* workspace,
                              size_t workspace_bytes, void* stream) {
    return litho_abbe_fft_accumulate_ex(p, maskFT, pupil, shifts, weights, n_src, batch, intensity, workspace,
                                        workspace_bytes, stream, 3);
}

static int accumulate_impl(const litho_plan_t* p, const void* maskFT, const void* pupil, int nf, size_t pupil_stride,
                           const int32_t* shifts, const float* weights, int n_src, int batch, float* intensity,
                           size_t intensity_stride, void* workspace, size_t workspace_bytes, void* stream, int phases);

int litho_abbe_fft_accumulate_ex(const litho_plan_t* p, const void* maskFT, const void* pupil, const int32_t* shifts,
                                 const float* weights, int n_src, int batch, float* intensity, void* workspace,
                                 size_t workspace_bytes, void* stream, int phases) {
    return accumulate_impl(p, maskFT, pupil, 1, 0, shifts, weights, n_src, batch, intensity, 0, workspace,
                           workspace_bytes, stream, phases);
}

int litho_abbe_fft_accumulate_focus(const litho_plan_t* p, const void* maskFT, const void* pupils, int n_focus,
                                    size_t pupil_stride, const int32_t* shifts, const float* weights, int n_src,
                                    int batch, float* intensities, size_t intensity_stride, void* workspace,
                                    size_t workspace_bytes, void* stream) {
    if (n_focus < 1 || n_focus > 4096) return fail(LITHO_ERR_ARG, "accumulate_focus: n_focus out of range");
    if (p && n_focus > 1) {
        if (pupil_stride < (size_t)p->pn * p->pn) return fail(LITHO_ERR_ARG, "accumulate_focus: pupil_stride < pn*pn");
        if (intensity_stride < intensity_elems(p)) return fail(LITHO_ERR_ARG, "accumulate_focus: intensity_stride < plan.intensity_elems");
    }
    return accumulate_impl(p, maskFT, pupils, n_focus, pupil_stride, shifts, weights, n_src, batch, intensities,
                           intensity_stride, workspace, workspace_bytes, stream, 3);
}

static int accumulate_impl(const litho_plan_t* p, const void* maskFT, const void* pupil, int nf, size_t pupil_stride,
                           const int32_t* shifts, const float* weights, int n_src, int batch, float* intensity,
                           size_t intensity_stride, void* workspace, size_t workspace_bytes, void* stream, int phases) {
    if (!p || !maskFT || !pupil || !intensity) return fail(LITHO_ERR_ARG, "accumulate: null argument");
    if (n_src < 0) return fail(LITHO_ERR_ARG, "accumulate: negative n_src");
    if (n_src == 0) return LITHO_OK;
    if (!shifts) return fail(LITHO_ERR_ARG, "accumulate: shifts is null");
    if (p->path != 2 && nf > 1) {
        // generic fine-grid kernels: one focus value at a time (same results, no sharing of the mask window)
        for (int f = 0; f < nf; ++f) {
            const int rc = accumulate_impl(p, maskFT, (const cplx*)pupil + (size_t)f * pupil_stride, 1, 0, shifts, weights,
                                           n_src, batch, intensity + (size_t)f * intensity_stride, 0, workspace,
                                           workspace_bytes, stream, phases);
            if (rc) return rc;
        }
        return LITHO_OK;
    }
    if (batch <= 0) {
        // default: the plan's batch, but at least 4 batches per call so that the row pass of one batch
        // overlaps the column pass of the previous one (short source lists, e.g. one rank's shard)
        batch = default_batch_focus(p, nf);
        const int q = (n_src + 3) / 4;
        if (q < batch) batch = q < 1 ? 1 : q;
    }
    if (batch > n_src) batch = n_src;
    if (!workspace || workspace_bytes < litho_plan_workspace_bytes_focus(p, batch, nf))
        return fail(LITHO_ERR_WORKSPACE, "accumulate: workspace too small for the requested batch");
    if ((double)batch * nf * 2.0 * p->Sr >= 2.0e9) return fail(LITHO_ERR_ARG, "accumulate: batch * n_focus too large");
    // equal-sized batches (130 points with batch 16 -> 9 x 14..15 instead of 8 x 16 + a 2-point launch pair)
    {
        const int nbatches = (n_src + batch - 1) / batch;
        batch = (n_src + nbatches - 1) / nbatches;
    }
    litho_stream_t st = (litho_stream_t)stream;
    if (p->path == 2) {
        // fast coarse-grid kernels; the caller guarantees shifts inside plan.shift_range (no wrap)
        FastRowsParams fr;
        memset(&fr, 0, sizeof(fr));
        fr.pupil = (const cplx*)pupil; fr.mask = (const cplx*)maskFT; fr.pn = p->pn;
        fr.pr0 = p->bbox[0]; fr.pc0 = p->bbox[2]; fr.Sr = p->Sr; fr.Sc = p->Sc;
        fr.shifts = (const int2_*)shifts; fr.tables = p->tables; fr.tables_c = p->tables_c; fr.tables_r = p->tables_r; fr.T = (cplx*)workspace;
        fr.status = p->status;
        fr.n_focus = nf; fr.pupil_stride = pupil_stride;
        FastColsParams fc;
        memset(&fc, 0, sizeof(fc));
        fc.T = (const cplx*)workspace; fc.Sr = p->Sr; fc.weights = weights; fc.tables = p->tables;
        fc.ic = intensity; fc.status = p->status;
        {
        // T is double-buffered: the row pass of batch b+1 runs on the plan's auxiliary stream while the
        // column pass of batch b runs on the caller's stream, so the tail of one kernel overlaps the head
        // of the other (both are short, ~20-40 us at cfg3).  rows(b) -> cols(b) and cols(b) -> rows(b+LITHO_TSLOTS)
        // (slot reuse) are ordered with events; all column passes stay on one stream, which also
        // serialises their read-modify-write of the intensity plane.
        const size_t focus_elems = (size_t)batch * 2 * p->Sr * p->Mf;   // T block of one focus value in a slot
        const size_t slot_elems = focus_elems * nf;
        fr.t_focus_stride = focus_elems;
        cplx* Tslot[LITHO_TSLOTS];
        for (int i = 0; i < LITHO_TSLOTS; ++i) Tslot[i] = (cplx*)workspace + (size_t)i * slot_elems;
        // one tensor map over the whole ring; a tile = Sr rows x tma_cols columns fetched in nbox boxes
        if (p->tma_cols > 0 && (phases & 2)) {
            // <= 4 boxes of M/4 rows cover the rows u < M that exist; the rim row u = M (Sr == M+1) is a 1-D copy
            const int body = p->Sr > p->Mf ? p->Mf : p->Sr;
            const int nbox = (body + p->tma_box_rows - 1) / p->tma_box_rows;
            if (make_tile_map(&fc.tile, workspace, p->Mf, (long long)LITHO_TSLOTS * nf * batch * 2 * p->Sr, p->tma_box_rows,
                              p->tma_cols) == 0) {
                fc.use_tma = p->tma_cols;
                fc.nbox = nbox;
                fc.rim = p->Sr > p->Mf ? p->Sr - p->Mf : 0;
                fc.tables_c = p->tables_c;
            }
        }
#if !defined(LITHO_EMU)
        const bool overlap = ((phases & 3) == 3) && p->aux_stream != nullptr;
        // LITHO_PHASE_INPUTS_READY: the row pass only reads the inputs and writes T, so it need not wait for
        // what the caller's stream is still doing (the previous image's last column pass, the zeroing of the
        // plane): its only dependencies are the column passes that last read each ring slot, which the plan's
        // events already track when the ring is the one the previous call used.
        const bool chained = overlap && (phases & LITHO_PHASE_INPUTS_READY) && p->last_ws == workspace &&
                             p->last_batch == batch * nf;
        if (overlap && !chained) {
            BE_CHECK((int)cudaEventRecord(p->ev_start, st));
            BE_CHECK((int)cudaStreamWaitEvent(p->aux_stream, p->ev_start, 0));
            for (int i = 0; i < LITHO_TSLOTS; ++i) p->cols_recorded[i] = 0;
        }
        p->last_ws = overlap ? workspace : nullptr;   // single-pass (profiling) calls break the chain
        p->last_batch = batch * nf;
#else
        const bool overlap = false;
#endif
        int b = 0;
        for (int s0 = 0; s0 < n_src; s0 += batch, ++b) {
            const int nb = (n_src - s0) < batch ? (n_src - s0) : batch;
            const int slot = b % LITHO_TSLOTS;
            fr.s_begin = s0; fr.batch = nb; fr.T = Tslot[slot];
            fc.s_begin = s0; fc.batch = nb;
            // one column pass per focus value: its T block is a contiguous [nb][2][Sr][M] batch, its plane its own
            auto cols_all = [&](litho_stream_t cs) -> int {
                for (int f = 0; f < nf; ++f) {
                    fc.T = Tslot[slot] + (size_t)f * focus_elems;
                    fc.row_begin = ((long long)slot * nf + f) * batch * 2 * p->Sr;
                    fc.ic = intensity + (size_t)f * intensity_stride;
                    const int rc = dispatch_fast_cols(p->Mf, p->ppt, fc, cs);
                    if (rc) return rc;
                }
                return 0;
            };
#if !defined(LITHO_EMU)
            if (overlap) {
                if (p->cols_recorded[slot])
                    BE_CHECK((int)cudaStreamWaitEvent(p->aux_stream, p->ev_cols[slot], 0));
                BE_CHECK(dispatch_fast_rows(p->Mf, p->ppt, fr, p->n_sm * (p->ppt == 16 ? 4 : 2), p->aux_stream));
                BE_CHECK((int)cudaEventRecord(p->ev_rows[slot], p->aux_stream));
                BE_CHECK((int)cudaStreamWaitEvent(st, p->ev_rows[slot], 0));
                BE_CHECK(cols_all(st));
                BE_CHECK((int)cudaEventRecord(p->ev_cols[slot], st));
                p->cols_recorded[slot] = 1;
                continue;
            }
#endif
            if (phases & 1) BE_CHECK(dispatch_fast_rows(p->Mf, p->ppt, fr, p->n_sm * (p->ppt == 16 ? 4 : 2), st));
            if (phases & 2) BE_CHECK(cols_all(st));
        }
        }
        if ((phases & 2) && p->q > 1 && (p->er >= 0 || p->ec >= 0))
        for (int f = 0; f < nf; ++f) {
            float* const intensity_f = intensity + (size_t)f * intensity_stride;
            RimParams rm;
            memset(&rm, 0, sizeof(rm));
            rm.pupil = (const cplx*)pupil + (size_t)f * pupil_stride; rm.mask = (const cplx*)maskFT; rm.pn = p->pn;
            rm.pr0 = p->bbox[0]; rm.pc0 = p->bbox[2]; rm.Sr = p->Sr; rm.Sc = p->Sc; rm.M = p->Mf;
            rm.shifts = (const int2_*)shifts; rm.weights = weights; rm.n_src = n_src;
            memcpy(rm.ext, p->ext, sizeof(rm.ext));
            rm.er = p->er; rm.ec = p->ec;
            // two stages, no atomics: LITHO_RIM_CHUNKS CTAs each sum a contiguous chunk of source points into a
            // private slice, then the slices are folded into the plane in chunk order (bit-reproducible)
            const int chunks = n_src < LITHO_RIM_CHUNKS ? n_src : LITHO_RIM_CHUNKS;
            rm.per_cta = (n_src + chunks - 1) / chunks;
            const int ctas = (n_src + rm.per_cta - 1) / rm.per_cta;
            rm.frow = p->rim_scratch; rm.fcol = p->rim_scratch + rim_row_floats(p); rm.stride = p->rim_stride;
            RimReduceParams rr;
            rr.slices = p->rim_scratch; rr.stride = p->rim_stride; rr.n_slices = ctas; rr.n = (int)p->rim_stride;
            rr.plane = rim_row_ptr(p, intensity_f);
            // shared memory: the longest (lo, hi) pair of lines that gets correlated
            int longest = 1;
            for (int axis = 0; axis < 2; ++axis) {
                const int e = axis == 0 ? p->er : p->ec;
                for (int k = 0; k <= e; ++k)
                    for (int t = 0; t + k <= e; ++t) {
                        const int b = e - k - t;
                        const int nl = p->ext[axis * 2][t][1] - p->ext[axis * 2][t][0] + 1;
                        const int nh = p->ext[axis * 2 + 1][b][1] - p->ext[axis * 2 + 1][b][0] + 1;
                        if (nl > 0 && nh > 0 && nl + nh > longest) longest = nl + nh;
                    }
            }
            const size_t smem = (size_t)longest * sizeof(cplx);
#if defined(LITHO_EMU)
            memset(p->rim_scratch, 0, (size_t)ctas * p->rim_stride * sizeof(float));
            litho_emu::launch(ctas, 1, 1, 64, smem, [&](const litho_emu::EmuCtx& c, char* s) { rim_body(rm, c, (cplx*)s); });
            for (int j = 0; j < rr.n; ++j) rim_reduce_elem(rr, j);
#else
            BE_CHECK((int)cudaMemsetAsync(p->rim_scratch, 0, (size_t)ctas * p->rim_stride * sizeof(float), st));
            if (smem > 48 * 1024)
                BE_CHECK((int)cudaFuncSetAttribute(rim_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            rim_kernel<<<ctas, 256, smem, st>>>(rm);
            BE_CHECK((int)cudaGetLastError());
            rim_reduce_kernel<<<(rr.n + 255) / 256, 256, 0, st>>>(rr);
            BE_CHECK((int)cudaGetLastError());
#endif
        }
        return LITHO_OK;
    }
    const int R = p->zp.R, M = p->zp.M;

    RowsParams rp;
    memset(&rp, 0, sizeof(rp));
    rp.pupil = (const cplx*)pupil; rp.mask = (const cplx*)maskFT; rp.pn = p->pn;
    rp.pr0 = p->bbox[0]; rp.pc0 = p->bbox[2];
    rp.shifts = (const int2_*)shifts;
    rp.lines = p->Sr;
    rp.ax = axis_in(p->bbox[2], p->pn, p->Sc);
    rp.out = p->out; rp.plan = p->zp; rp.twL = p->twL; rp.T = (cplx*)workspace;

    ColsParams cp;
    memset(&cp, 0, sizeof(cp));
    cp.T = (const cplx*)workspace;
    cp.Rc = R; cp.Wrc = p->zp.Wr; cp.outc = p->out;
    cp.shifts = (const int2_*)shifts; cp.weights = weights;
    cp.ax = axis_in(p->bbox[0], p->pn, p->Sr);
    cp.out = p->out; cp.plan = p->zp; cp.twL = p->twL;
    cp.iperm = intensity; cp.scale = 1.f;

    const int gx_rows = (p->Sr * R + p->rows_fpc - 1) / p->rows_fpc;
    const int gx_cols = R * ((p->zp.Wr + p->cols_cb - 1) / p->cols_cb);
    for (int s0 = 0; s0 < n_src; s0 += batch) {
        const int nb = (n_src - s0) < batch ? (n_src - s0) : batch;
        rp.s_begin = s0;
        if (phases & 1) BE_CHECK(dispatch_rows(M, ROW_PUPIL_MASK, rp, gx_rows, nb, st));
        cp.s_begin = s0; cp.batch = nb;
        if (phases & 2) BE_CHECK(dispatch_cols(M, EPI_ACCUM, cp, gx_cols, R, st));
    }
    return LITHO_OK;
}

static PermView perm_view(const litho_plan* p, const float* intensity) {
    PermView v;
    memset(&v, 0, sizeof(v));
    v.iperm = intensity;
    v.outr = p->out; v.outc = p->out;
    v.Rr = p->zp.R; v.Rc = p->zp.R; v.Wrr = p->zp.Wr; v.Wrc = p->zp.Wr;
    if (p->path == 2) {  // q == 1: the coarse grid is the fine grid, read it in place
        v.mode = 2; v.Mc = p->Mf; v.center = p->pn / 2;
    }
    return v;
}

// Fast path with q > 1: exact (spectral) interpolation of the coarse intensity to the pn centre pixels
// of the reference's N grid.  Leaves the natural pn x pn plane at ws + L.fine_off and returns a view on it.
//   1. coarse plane -> natural order                      (coarse_unperm)
//   2. centred forward DFT, scaled 1/Nc^2                 (generic zoom kernels, conjugating epilogue)
//   3. (Nc+1)^2 spectrum with the +-Mf lines from the rim sums   (assemble)
//   4. inverse zoom DFT of length N to the pn centre pixels, real part   (generic zoom kernels)
static int fine_plane_fast(const litho_plan* p, const float* intensity, void* workspace, size_t workspace_bytes,
                           litho_stream_t st, PermView* view) {
    FinalizeLayout L;
    if (finalize_layout(p, &L)) return fail(LITHO_ERR_ARG, "finalize: unsupported coarse grid");
    if (!workspace || workspace_bytes < L.total) return fail(LITHO_ERR_WORKSPACE, "finalize: workspace too small");
    char* ws = (char*)workspace;
    float* A = (float*)(ws + L.a_off);
    cplx* T1 = (cplx*)(ws + L.t1_off);
    cplx* fhat = (cplx*)(ws + L.fhat_off);
    cplx* spec = (cplx*)(ws + L.spec_off);
    cplx* T2 = (cplx*)(ws + L.t2_off);
    float* fine = (float*)(ws + L.fine_off);
    const int Nc = p->Nc, pn = p->pn, N = p->N;
    cplx *tw1 = nullptr, *tw2 = nullptr;
    BE_CHECK(get_twiddles(Nc, &tw1));
    BE_CHECK(get_twiddles(N, &tw2));

    CoarseUnpermParams cu{intensity, p->Mf, A};
#if defined(LITHO_EMU)
    for (int a = 0; a < Nc; ++a)
        for (int b = 0; b < Nc; ++b) coarse_unperm_elem(cu, a, b);
#else
    coarse_unperm_kernel<<<dim3((Nc + 255) / 256, Nc, 1), 256, 0, st>>>(cu);
    BE_CHECK((int)cudaGetLastError());
#endif
    const int BIG = 1 << 30;
    {   // step 2: Fhat[m][n] = (1/Nc^2) sum_ab A[a][b] exp(-2 pi i (m a + n b)/Nc), m,n in [-Nc/2, Nc/2)
        AxisIn ax; ax.first = 0; ax.period = BIG; ax.center = 0; ax.S = Nc;
        RowsParams rp; memset(&rp, 0, sizeof(rp));
        rp.real_in = A; rp.in_pitch = Nc; rp.lines = Nc; rp.ax = ax; rp.out = L.o1; rp.plan = L.z1; rp.twL = tw1; rp.T = T1;
        ColsParams cp; memset(&cp, 0, sizeof(cp));
        cp.T = T1; cp.batch = 1; cp.Rc = 1; cp.Wrc = Nc; cp.outc = L.o1; cp.ax = ax; cp.out = L.o1; cp.plan = L.z1;
        cp.twL = tw1; cp.field = fhat; cp.field_pitch = Nc; cp.conj_out = 1;
        cp.scale = (float)(1.0 / ((double)Nc * (double)Nc));
        BE_CHECK(dispatch_rows(Nc, ROW_REAL_PLANE, rp, (Nc + L.fpc1 - 1) / L.fpc1, 1, st));
        BE_CHECK(dispatch_cols(Nc, EPI_FIELD, cp, (Nc + L.cb1 - 1) / L.cb1, 1, st));
    }
    AssembleParams as;
    as.fhat = fhat; as.frow = rim_row_ptr(p, const_cast<float*>(intensity));
    as.fcol = rim_col_ptr(p, const_cast<float*>(intensity));
    as.Nc = Nc; as.er = p->er; as.ec = p->ec; as.Sr = p->Sr; as.Sc = p->Sc; as.out = spec;
    const int SR = L.spec_rows, SC = L.spec_cols;
#if defined(LITHO_EMU)
    for (int i = 0; i < SR; ++i)
        for (int j = 0; j < SC; ++j) assemble_elem(as, i, j);
#else
    assemble_kernel<<<dim3((SC + 255) / 256, SR, 1), 256, 0, st>>>(as, SC);
    BE_CHECK((int)cudaGetLastError());
#endif
    {   // step 4: fine[i][j] = Re sum_{m,n} spec[m][n] exp(+2 pi i (m i' + n j')/N), i' = i - pn/2
        AxisIn axc; axc.first = 0; axc.period = BIG; axc.center = (SC - 1) / 2; axc.S = SC;   // along a spectrum row
        AxisIn axr; axr.first = 0; axr.period = BIG; axr.center = (SR - 1) / 2; axr.S = SR;   // along a spectrum column
        const int R = L.z2.R, M2 = L.z2.M;
        RowsParams rp; memset(&rp, 0, sizeof(rp));
        rp.cplx_in = spec; rp.in_pitch = SC; rp.lines = SR; rp.ax = axc; rp.out = L.o2; rp.plan = L.z2;
        rp.twL = tw2; rp.T = T2;
        ColsParams cp; memset(&cp, 0, sizeof(cp));
        cp.T = T2; cp.batch = 1; cp.Rc = R; cp.Wrc = L.z2.Wr; cp.outc = L.o2; cp.ax = axr; cp.out = L.o2; cp.plan = L.z2;
        cp.twL = tw2; cp.real_out = fine; cp.field_pitch = pn; cp.conj_out = 0; cp.scale = 1.f;
        BE_CHECK(dispatch_rows(M2, ROW_CPLX_PLANE, rp, (SR * R + L.fpc2 - 1) / L.fpc2, 1, st));
        BE_CHECK(dispatch_cols(M2, EPI_FIELD, cp, R * ((L.z2.Wr + L.cb2 - 1) / L.cb2), R, st));
    }
    memset(view, 0, sizeof(*view));
    view->mode = 1; view->iperm = fine; view->pitch = pn;
    return LITHO_OK;
}

static int source_view(const litho_plan* p, const float* intensity, void* workspace, size_t workspace_bytes,
                       litho_stream_t st, PermView* view) {
    if (p->path == 2 && p->q > 1) return fine_plane_fast(p, intensity, workspace, workspace_bytes, st, view);
    *view = perm_view(p, intensity);
    return LITHO_OK;
}

// imageformation.py:71-75 size arithmetic (python semantics: floor, round-half-even, floor division)
static void post_sizes(int pn, double eps, int* side, int* pW, int* out_side) {
    const double sf = 1.0 / eps;
    *side = (int)floor((double)pn * sf);
    const long rnd = (long)nearbyint((double)pn / eps);
    const long diff = (long)pn - rnd;
    *pW = (int)(diff >= 0 ? diff / 2 : -((-diff + 1) / 2));
    const int corr = *side % 2;
    *out_side = *side + 2 * (*pW) + corr;
}

int litho_fft_output_side(int pn, double eps) {
    int side, pW, os;
    post_sizes(pn, eps, &side, &pW, &os);
    return os;
}

int litho_abbe_fft_finalize(const litho_plan_t* p, const float* intensity, double eps, float* out, void* workspace,
                            size_t workspace_bytes, void* stream) {
    if (!p || !intensity || !out || !(eps > 0)) return fail(LITHO_ERR_ARG, "finalize: bad argument");
    FinalizeParams F;
    int rcv = source_view(p, intensity, workspace, workspace_bytes, (litho_stream_t)stream, &F.in);
    if (rcv) return rcv;
    F.pn = p->pn;
    post_sizes(p->pn, eps, &F.side, &F.pW, &F.out_side);
    if (F.out_side <= 0) return fail(LITHO_ERR_ARG, "finalize: empty output");
    F.scale = (float)(1.0 / (1.0 / eps));
    F.out = out;
#if defined(LITHO_EMU)
    (void)stream;
    for (int y = 0; y < F.out_side; ++y)
        for (int x = 0; x < F.out_side; ++x) finalize_pixel(F, y, x);
#else
    dim3 grid((F.out_side + 255) / 256, F.out_side, 1);
    finalize_kernel<<<grid, 256, 0, (litho_stream_t)stream>>>(F);
    BE_CHECK((int)cudaGetLastError());
#endif
    return LITHO_OK;
}

int litho_abbe_fft_unpermute(const litho_plan_t* p, const float* intensity, float* out, void* workspace,
                             size_t workspace_bytes, void* stream) {
    if (!p || !intensity || !out) return fail(LITHO_ERR_ARG, "unpermute: null argument");
    UnpermParams U;
    int rcv = source_view(p, intensity, workspace, workspace_bytes, (litho_stream_t)stream, &U.in);
    if (rcv) return rcv;
    U.pn = p->pn;
    U.out = out;
#if defined(LITHO_EMU)
    (void)stream;
    for (int y = 0; y < U.pn; ++y)
        for (int x = 0; x < U.pn; ++x) out[(size_t)y * U.pn + x] = U.in.at(y, x);
#else
    dim3 grid((U.pn + 255) / 256, U.pn, 1);
    unpermute_kernel<<<grid, 256, 0, (litho_stream_t)stream>>>(U);
    BE_CHECK((int)cudaGetLastError());
#endif
    return LITHO_OK;
}

int litho_fft_field(const litho_plan_t* p, const void* pf, const void* maskFT, void* field, void* workspace,
                    size_t workspace_bytes, void* stream) {
    if (!p || !pf || !maskFT || !field) return fail(LITHO_ERR_ARG, "fft_field: null argument");
    if (!workspace || workspace_bytes < litho_plan_workspace_bytes(p, 1))
        return fail(LITHO_ERR_WORKSPACE, "fft_field: workspace too small");
    litho_stream_t st = (litho_stream_t)stream;
    const int R = p->zp.R, M = p->zp.M;
    RowsParams rp;
    memset(&rp, 0, sizeof(rp));
    rp.pupil = (const cplx*)pf; rp.mask = (const cplx*)maskFT; rp.pn = p->pn;
    rp.pr0 = p->bbox[0]; rp.pc0 = p->bbox[2];
    rp.shifts = nullptr;
    rp.lines = p->Sr;
    rp.ax = axis_in(p->bbox[2], p->pn, p->Sc);
    rp.out = p->out; rp.plan = p->zp; rp.twL = p->twL; rp.T = (cplx*)workspace;
    ColsParams cp;
    memset(&cp, 0, sizeof(cp));
    cp.T = (const cplx*)workspace;
    cp.batch = 1;
    cp.Rc = R; cp.Wrc = p->zp.Wr; cp.outc = p->out;
    cp.ax = axis_in(p->bbox[0], p->pn, p->Sr);
    cp.out = p->out; cp.plan = p->zp; cp.twL = p->twL;
    cp.field = (cplx*)field; cp.field_pitch = p->pn; cp.conj_out = 0; cp.scale = 1.f;
    const int gx_rows = (p->Sr * R + p->rows_fpc - 1) / p->rows_fpc;
    const int gx_cols = R * ((p->zp.Wr + p->cols_cb - 1) / p->cols_cb);
    BE_CHECK(dispatch_rows(M, ROW_PUPIL_MASK, rp, gx_rows, 1, st));
    BE_CHECK(dispatch_cols(M, EPI_FIELD, cp, gx_cols, R, st));
    return LITHO_OK;
}

// ---------------------------------------------------------------------------- mask spectrum
// Mask._ffFraunhofer (mask.py:74-90): bilinear upsample of the int16 geometry by eps, centre-pad
// to N, centred forward DFT, crop to pn.  The forward transform of a real plane is the conjugate
// of the inverse one, so the imaging kernels are reused with a conjugating epilogue.
struct SpecGeom {
    int pn, N, sm;      // grid, transform length, resampled side
    int u0, S, first;   // first resampled sample used, how many, and its position in the N grid
    ZoomPlan zp;
    AxisOut out;
    int rows_fpc, cols_cb;
    size_t plane_elems, t_elems;
};

static int spec_geom(int pn, double eps, int N, SpecGeom* g) {
    if (pn < 2 || (pn & 1)) return 1;
    if (!is_pow2(N) || N > 16384 || N < 16 || N < pn) return 1;
    if (!(eps > 0)) return 1;
    g->pn = pn; g->N = N;
    g->sm = (int)floor((double)pn * eps);
    if (g->sm < 1) return 1;
    const long d = (long)(N - pn) - (long)(g->sm - pn);
    const long pW = d >= 0 ? d / 2 : -((-d + 1) / 2);  // python floor division
    g->u0 = pW < 0 ? (int)(-pW) : 0;                   // negative pad = crop (F.pad semantics)
    g->first = pW < 0 ? 0 : (int)pW;
    int S = g->sm - g->u0;
    if (S > N - g->first) S = N - g->first;
    if (S < 1) return 1;
    g->S = S;
    int M = 16;
    while (M < S - 1) M <<= 1;
    if (M > N) M = N;
    g->zp.L = N; g->zp.M = M; g->zp.R = N / M;
    g->out.W = pn; g->out.center = pn / 2;
    g->zp.Wr = (pn + g->zp.R - 1) / g->zp.R;
    if (dispatch_shape(M, &g->rows_fpc, &g->cols_cb)) return 1;
    g->plane_elems = (size_t)g->sm * g->sm;
    g->t_elems = (size_t)g->zp.R * S * g->zp.Wr;
    return 0;
}

size_t litho_mask_spectrum_workspace_bytes(int pn, double eps, int N) {
    SpecGeom g;
    if (spec_geom(pn, eps, N, &g)) return 0;
    return ((g.plane_elems * sizeof(float) + 15) / 16) * 16 + g.t_elems * sizeof(cplx);
}

int litho_mask_spectrum(const int16_t* geometry, int pn, double eps, int N, void* maskFT, void* workspace,
                        size_t workspace_bytes, void* stream) {
    if (!geometry || !maskFT || !workspace) return fail(LITHO_ERR_ARG, "mask_spectrum: null argument");
    SpecGeom g;
    if (spec_geom(pn, eps, N, &g)) return fail(LITHO_ERR_ARG, "mask_spectrum: unsupported pn / eps / N");
    if (workspace_bytes < litho_mask_spectrum_workspace_bytes(pn, eps, N))
        return fail(LITHO_ERR_WORKSPACE, "mask_spectrum: workspace too small");
    litho_stream_t st = (litho_stream_t)stream;
    float* plane = (float*)workspace;
    cplx* T = (cplx*)((char*)workspace + ((g.plane_elems * sizeof(float) + 15) / 16) * 16);
    cplx* tw = nullptr;
    BE_CHECK(get_twiddles(N, &tw));

    ResampleParams rs;
    rs.in = geometry; rs.pn = pn; rs.side = g.sm; rs.scale = (float)(1.0 / eps); rs.out = plane;
#if defined(LITHO_EMU)
    for (int y = 0; y < g.sm; ++y)
        for (int x = 0; x < g.sm; ++x) resample_pixel(rs, y, x);
#else
    {
        dim3 grid((g.sm + 255) / 256, g.sm, 1);
        resample_kernel<<<grid, 256, 0, st>>>(rs);
        BE_CHECK((int)cudaGetLastError());
    }
#endif
    const int R = g.zp.R, M = g.zp.M;
    AxisIn ax;
    ax.first = g.first; ax.period = 1 << 30; ax.center = N / 2; ax.S = g.S;

    RowsParams rp;
    memset(&rp, 0, sizeof(rp));
    rp.real_in = plane + (size_t)g.u0 * g.sm + g.u0;
    rp.in_pitch = g.sm;
    rp.lines = g.S;
    rp.ax = ax; rp.out = g.out; rp.plan = g.zp; rp.twL = tw; rp.T = T;
    ColsParams cp;
    memset(&cp, 0, sizeof(cp));
    cp.T = T; cp.batch = 1; cp.Rc = R; cp.Wrc = g.zp.Wr; cp.outc = g.out;
    cp.ax = ax; cp.out = g.out; cp.plan = g.zp; cp.twL = tw;
    cp.field = (cplx*)maskFT; cp.field_pitch = pn; cp.conj_out = 1; cp.scale = 1.f;
    const int gx_rows = (g.S * R + g.rows_fpc - 1) / g.rows_fpc;
    const int gx_cols = R * ((g.zp.Wr + g.cols_cb - 1) / g.cols_cb);
    BE_CHECK(dispatch_rows(M, ROW_REAL_PLANE, rp, gx_rows, 1, st));
    BE_CHECK(dispatch_cols(M, EPI_FIELD, cp, gx_cols, R, st));
    return LITHO_OK;
}

// ---------------------------------------------------------------------------- direct solver
static int direct_check(int pn, const int* bbox, int* Sr, int* Sc) {
    if (pn < 2 || !bbox) return 1;
    if (bbox[0] < 0 || bbox[2] < 0 || bbox[1] >= pn || bbox[3] >= pn || bbox[1] < bbox[0] || bbox[3] < bbox[2]) return 1;
    *Sr = bbox[1] - bbox[0] + 1;
    *Sc = bbox[3] - bbox[2] + 1;
    return 0;
}

int litho_direct_operator(int pn, double pixelSize, double wavelength, int sign, void* A, void* stream) {
    if (!A || pn < 2 || !(pixelSize > 0) || !(wavelength > 0) || (sign != 1 && sign != -1))
        return fail(LITHO_ERR_ARG, "direct_operator: bad argument");
    // python-double arithmetic of imageformation.py:4-8 / mask.py:32-35, then float32 like torch.arange
    const double deltaK = 4.0 / (double)pn;
    const double Kbound = (double)pn / 2.0 * deltaK;
    const double pixelBound = (double)pn / 2.0 * pixelSize;
    DirectOpParams P;
    P.kstart = (float)(-Kbound); P.kstep = (float)deltaK;
    P.xstart = (float)(-pixelBound); P.xstep = (float)pixelSize;
    P.c0 = (float)(2.0 * M_PI / wavelength);
    P.sign = sign; P.pn = pn; P.A = (cplx*)A;
#if defined(LITHO_EMU)
    (void)stream;
    for (int a = 0; a < pn; ++a)
        for (int c = 0; c < pn; ++c) direct_op_elem(P, a, c);
#else
    dim3 grid((pn + 255) / 256, pn, 1);
    direct_op_kernel<<<grid, 256, 0, (litho_stream_t)stream>>>(P);
    BE_CHECK((int)cudaGetLastError());
#endif
    return LITHO_OK;
}

// Tensor-core path of litho_direct_accumulate (direct_tc.cu): on by default in the device build, LITHO_DIRECT_TC=0
// selects the FP32 CUDA-core kernels (also used by the field / mask-spectrum entry points and the CPU emulation).
static bool direct_tc_enabled(int pn) {
#if defined(LITHO_EMU)
    (void)pn;
    return false;
#else
    if (pn < 64) return false;
    const char* env = getenv("LITHO_DIRECT_TC");
    return !(env && atoi(env) == 0);
#endif
}
static int tc_upitch(int Sr) { return (Sr + 15) / 16 * 16; }

// Source points per launch group when the caller passes batch <= 0.  FP32 kernels: 8.  Tensor-core kernels: one CTA
// covers a 128-row x <=128-column tile of ONE source point, so enough points to put two CTAs on every SM (capped at
// 128 points: the workspace holds U and |E|^2 per point).
int litho_direct_default_batch(int pn, const int* bbox) {
    int Sr, Sc;
    if (direct_check(pn, bbox, &Sr, &Sc)) return 0;
#if !defined(LITHO_EMU)
    if (direct_tc_enabled(pn)) {
        int sms = 148, dev = 0;
        if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        const int per_point = litho_tc::tc_ctas_per_point(pn, Sr);    // stage 1 has the fewer CTAs per point
        int b = (2 * sms + per_point - 1) / per_point;
        return b < 8 ? 8 : (b > 128 ? 128 : b);
    }
#endif
    return 8;
}

size_t litho_direct_workspace_bytes(int pn, const int* bbox, int batch) {
    int Sr, Sc;
    if (direct_check(pn, bbox, &Sr, &Sc)) return 0;
    if (batch < 1) batch = litho_direct_default_batch(pn, bbox);
    if (batch < 1) return 0;
    const size_t fp32 = (size_t)batch * Sr * pn * sizeof(cplx);
    const size_t tc = (size_t)batch * pn * tc_upitch(Sr) * sizeof(cplx) + (size_t)batch * pn * pn * sizeof(float);
    return (direct_tc_enabled(pn) && tc > fp32) ? tc : fp32;
}

#if !defined(LITHO_EMU)
static std::mutex g_tc_err_mutex;
static std::map<int, int*> g_tc_err;
static int* tc_err_word() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
    std::lock_guard<std::mutex> lock(g_tc_err_mutex);
    auto it = g_tc_err.find(dev);
    if (it != g_tc_err.end()) return it->second;
    int* d = nullptr;
    if (cudaMalloc((void**)&d, sizeof(int)) != cudaSuccess) return nullptr;
    cudaMemset(d, 0, sizeof(int));
    g_tc_err[dev] = d;
    return d;
}
#endif

int litho_direct_status(int* status_host, void* stream) {
    if (!status_host) return fail(LITHO_ERR_ARG, "direct_status: null argument");
    *status_host = 0;
#if !defined(LITHO_EMU)
    int* d = tc_err_word();
    if (!d) return fail(LITHO_ERR_CUDA, "direct_status: no error word");
    BE_CHECK(be_d2h_sync(status_host, d, sizeof(int), (litho_stream_t)stream));
    if (*status_host) {
        const int zero = 0;
        BE_CHECK(be_h2d(d, &zero, sizeof(int), (litho_stream_t)stream));
        BE_CHECK((int)cudaStreamSynchronize((litho_stream_t)stream));
    }
#else
    (void)stream;
#endif
    return LITHO_OK;
}

static int direct_launch(int kind, int epi, const DirectParams& P, litho_stream_t st) {
    const int gb = (P.pn + DT - 1) / DT, gu = (P.Sr + DT - 1) / DT;
    const size_t smem = DIRECT_SMEM_ELEMS * sizeof(cplx);
#if defined(LITHO_EMU)
    (void)st;
    if (kind == DIRECT_PUPIL_MASK)
        litho_emu::launch(gb, gu, P.batch, DIRECT_THREADS, smem,
                          [&](const litho_emu::EmuCtx& c, char* s) { direct_rows_body<DIRECT_PUPIL_MASK>(P, c, (cplx*)s); });
    else
        litho_emu::launch(gb, gu, P.batch, DIRECT_THREADS, smem,
                          [&](const litho_emu::EmuCtx& c, char* s) { direct_rows_body<DIRECT_GEOMETRY>(P, c, (cplx*)s); });
    if (epi == DIRECT_ACCUM)
        litho_emu::launch(gb, gb, 1, DIRECT_THREADS, smem,
                          [&](const litho_emu::EmuCtx& c, char* s) { direct_cols_body<DIRECT_ACCUM>(P, c, (cplx*)s); });
    else
        litho_emu::launch(gb, gb, 1, DIRECT_THREADS, smem,
                          [&](const litho_emu::EmuCtx& c, char* s) { direct_cols_body<DIRECT_FIELD>(P, c, (cplx*)s); });
    return 0;
#else
    dim3 g1(gb, gu, P.batch), g2(gb, gb, 1);
    if (kind == DIRECT_PUPIL_MASK) direct_rows_kernel<DIRECT_PUPIL_MASK><<<g1, DIRECT_THREADS, smem, st>>>(P);
    else direct_rows_kernel<DIRECT_GEOMETRY><<<g1, DIRECT_THREADS, smem, st>>>(P);
    int e = (int)cudaGetLastError();
    if (e) return e;
    if (epi == DIRECT_ACCUM) direct_cols_kernel<DIRECT_ACCUM><<<g2, DIRECT_THREADS, smem, st>>>(P);
    else direct_cols_kernel<DIRECT_FIELD><<<g2, DIRECT_THREADS, smem, st>>>(P);
    return (int)cudaGetLastError();
#endif
}

int litho_direct_accumulate(const void* A, const void* maskFT, const void* pupil, int pn, const int* bbox,
                            const int32_t* shifts, const float* weights, int n_src, int batch, float* intensity,
                            void* workspace, size_t workspace_bytes, void* stream) {
    int Sr, Sc;
    if (!A || !maskFT || !pupil || !intensity || direct_check(pn, bbox, &Sr, &Sc))
        return fail(LITHO_ERR_ARG, "direct_accumulate: bad argument");
    if (n_src < 0) return fail(LITHO_ERR_ARG, "direct_accumulate: negative n_src");
    if (n_src == 0) return LITHO_OK;
    if (!shifts) return fail(LITHO_ERR_ARG, "direct_accumulate: shifts is null");
    if (batch < 1) batch = litho_direct_default_batch(pn, bbox);
    if (batch > n_src) batch = n_src;
    if (!workspace || workspace_bytes < litho_direct_workspace_bytes(pn, bbox, batch))
        return fail(LITHO_ERR_WORKSPACE, "direct_accumulate: workspace too small");
#if !defined(LITHO_EMU)
    if (direct_tc_enabled(pn)) {
        // tensor cores: both products as 3xTF32 real GEMMs with fp32 accumulation in tensor memory (direct_tc.cu)
        litho_tc::TcParams T;
        memset(&T, 0, sizeof(T));
        T.A = (const float2*)A; T.pn = pn; T.pupil = (const float2*)pupil; T.mask = (const float2*)maskFT;
        T.pr0 = bbox[0]; T.pc0 = bbox[2]; T.Sr = Sr; T.Sc = Sc;
        T.shifts = (const int2*)shifts;
        T.Upitch = tc_upitch(Sr);
        float2* U = (float2*)workspace;
        float* part = (float*)((char*)workspace + (size_t)batch * pn * T.Upitch * sizeof(cplx));
        T.U = U; T.Uout = U; T.part = part; T.err = tc_err_word();
        for (int s0 = 0; s0 < n_src; s0 += batch) {
            const int nb = (n_src - s0) < batch ? (n_src - s0) : batch;
            T.s_begin = s0;
            T.stage = 1;
            BE_CHECK(litho_tc::tc_launch(T, Sr, nb, (litho_stream_t)stream));
            T.stage = 2;
            BE_CHECK(litho_tc::tc_launch(T, pn, nb, (litho_stream_t)stream));
            BE_CHECK(litho_tc::tc_reduce(part, weights, s0, nb, pn, intensity, (litho_stream_t)stream));
        }
        return LITHO_OK;
    }
#endif
    DirectParams P;
    memset(&P, 0, sizeof(P));
    P.A = (const cplx*)A; P.pn = pn; P.pupil = (const cplx*)pupil; P.mask = (const cplx*)maskFT;
    P.pr0 = bbox[0]; P.pc0 = bbox[2]; P.Sr = Sr; P.Sc = Sc;
    P.shifts = (const int2_*)shifts; P.weights = weights;
    P.T = (cplx*)workspace; P.intensity = intensity;
    for (int s0 = 0; s0 < n_src; s0 += batch) {
        P.s_begin = s0;
        P.batch = (n_src - s0) < batch ? (n_src - s0) : batch;
        BE_CHECK(direct_launch(DIRECT_PUPIL_MASK, DIRECT_ACCUM, P, (litho_stream_t)stream));
    }
    return LITHO_OK;
}

int litho_direct_field(const void* A, const void* pupil, const void* maskFT, int pn, const int* bbox, void* field,
                       void* workspace, size_t workspace_bytes, void* stream) {
    int Sr, Sc;
    if (!A || !maskFT || !pupil || !field || direct_check(pn, bbox, &Sr, &Sc))
        return fail(LITHO_ERR_ARG, "direct_field: bad argument");
    if (!workspace || workspace_bytes < litho_direct_workspace_bytes(pn, bbox, 1))
        return fail(LITHO_ERR_WORKSPACE, "direct_field: workspace too small");
    DirectParams P;
    memset(&P, 0, sizeof(P));
    P.A = (const cplx*)A; P.pn = pn; P.pupil = (const cplx*)pupil; P.mask = (const cplx*)maskFT;
    P.pr0 = bbox[0]; P.pc0 = bbox[2]; P.Sr = Sr; P.Sc = Sc;
    P.batch = 1; P.T = (cplx*)workspace; P.field = (cplx*)field;
    BE_CHECK(direct_launch(DIRECT_PUPIL_MASK, DIRECT_FIELD, P, (litho_stream_t)stream));
    return LITHO_OK;
}

int litho_direct_mask_spectrum(const void* Aplus, const int16_t* geometry, int pn, void* maskFT, void* workspace,
                               size_t workspace_bytes, void* stream) {
    if (!Aplus || !geometry || !maskFT || pn < 2) return fail(LITHO_ERR_ARG, "direct_mask_spectrum: bad argument");
    const int bbox[4] = {0, pn - 1, 0, pn - 1};
    if (!workspace || workspace_bytes < litho_direct_workspace_bytes(pn, bbox, 1))
        return fail(LITHO_ERR_WORKSPACE, "direct_mask_spectrum: workspace too small");
    DirectParams P;
    memset(&P, 0, sizeof(P));
    P.A = (const cplx*)Aplus; P.pn = pn; P.geometry = geometry;
    P.pr0 = 0; P.pc0 = 0; P.Sr = pn; P.Sc = pn;
    P.batch = 1; P.T = (cplx*)workspace; P.field = (cplx*)maskFT;
    BE_CHECK(direct_launch(DIRECT_GEOMETRY, DIRECT_FIELD, P, (litho_stream_t)stream));
    return LITHO_OK;
}

// ---------------------------------------------------------------------------- builders
int litho_source_build(int pn, double sigma_in, double sigma_out, double shift_x, double shift_y, int quasar_count,
                       double rotation, int64_t* out, void* stream) {
    if (!out || pn < 1 || quasar_count < 0 || quasar_count > 16)
        return fail(LITHO_ERR_ARG, "source_build: bad argument (quasar_count must be 0..16)");
    SourceParams P;
    memset(&P, 0, sizeof(P));
    P.pn = pn;
    const double span = 2.0;
    P.x_start = (float)(-span - shift_x);
    P.y_start = (float)(-span - shift_y);
    P.step = (float)(span * 2.0 / (double)pn);
    P.sigma_in = round_f16((float)sigma_in);
    P.sigma_out = round_f16((float)sigma_out);
    P.quasar = quasar_count > 0;
    P.count = quasar_count;
    P.rotation = (float)rotation;
    P.two_pi = round_f16((float)(2.0 * M_PI));
    if (quasar_count > 0) {
        const double spacing = M_PI / (double)quasar_count;
        for (int g = 0; g < quasar_count; ++g) {
            P.spacing_lo[g] = round_f16((float)((double)(g + g) * spacing));
            P.spacing_hi[g] = round_f16((float)((double)(g + g + 1) * spacing));
        }
    }
    P.out = out;
#if defined(LITHO_EMU)
    (void)stream;
    for (int i = 0; i < pn; ++i)
        for (int j = 0; j < pn; ++j) source_pixel(P, i, j);
#else
    source_kernel<<<dim3((pn + 255) / 256, pn, 1), 256, 0, (litho_stream_t)stream>>>(P);
    BE_CHECK((int)cudaGetLastError());
#endif
    return LITHO_OK;
}

static double factorial_d(int n) {
    double f = 1.0;
    for (int i = 2; i <= n; ++i) f *= (double)i;
    return f;
}

int litho_pupil_build(const float* aberrations_host, int n_ab, int pn, void* pupil, void* wavefront, void* stream) {
    if (!aberrations_host || n_ab < 1 || n_ab > 256 || pn < 1 || (!pupil && !wavefront))
        return fail(LITHO_ERR_ARG, "pupil_build: bad argument");
    std::vector<ZernikeTerm> terms(n_ab);
    for (int j = 0; j < n_ab; ++j) {
        ZernikeTerm& z = terms[j];
        memset(&z, 0, sizeof(z));
        // OSA/ANSI index -> (m, n), pupil.py:82-86
        const int n = (int)ceil(0.5 * (-3.0 + sqrt(9.0 + 8.0 * (double)j)));
        const int m = 2 * j - n * (n + 2);
        const int am = m < 0 ? -m : m;
        const int lo = (n - am) / 2, hi = (n + am) / 2;
        if (lo + 1 > 8) return fail(LITHO_ERR_ARG, "pupil_build: radial order too high (n <= 15 supported)");
        z.m = m; z.n = n; z.nk = lo + 1;
        for (int k = 0; k <= lo; ++k) {
            const double c = ((k & 1) ? -1.0 : 1.0) * factorial_d(n - k) /
                             (factorial_d(k) * factorial_d(hi - k) * factorial_d(lo - k));
            z.stat[k] = (float)c;
            z.expo[k] = n - 2 * k;
        }
        const double nmn = sqrt((2.0 * n + 1.0) / (1.0 + (m == 0 ? 1.0 : 0.0)));
        const float coeff = round_f16(aberrations_host[j]);
        z.cn = round_f16(coeff * (float)(m >= 0 ? nmn : -nmn));
    }
    litho_stream_t st = (litho_stream_t)stream;
    ZernikeTerm* dterms = nullptr;
    std::unique_lock<std::mutex> scratch_lock(g_scratch_mutex, std::defer_lock);
    const bool own_terms = sizeof(ZernikeTerm) * (size_t)n_ab > SCRATCH_BYTES;   // > ~200 terms: a buffer of its own
    if (own_terms) {
        BE_CHECK(be_malloc((void**)&dterms, sizeof(ZernikeTerm) * n_ab));
    } else {
        scratch_lock.lock();
        int* sc = nullptr;
        BE_CHECK(get_scratch(&sc));
        dterms = reinterpret_cast<ZernikeTerm*>(sc);
    }
    int rc = be_h2d(dterms, terms.data(), sizeof(ZernikeTerm) * n_ab, st);
    PupilParams P;
    memset(&P, 0, sizeof(P));
    P.pn = pn; P.start = -2.0f; P.step = (float)(4.0 / (double)pn);
    P.n_terms = n_ab; P.terms = dterms;
    P.two_pi_f = (float)(2.0 * M_PI);
    P.pupil = (cplx*)pupil; P.we = (cplx*)wavefront;
    if (rc == 0) {
#if defined(LITHO_EMU)
        for (int i = 0; i < pn; ++i)
            for (int j = 0; j < pn; ++j) pupil_pixel(P, i, j);
#else
        pupil_kernel<<<dim3((pn + 255) / 256, pn, 1), 256, 0, st>>>(P);
        rc = (int)cudaGetLastError();
        if (rc == 0) rc = (int)cudaStreamSynchronize(st);  // the term table is freed below
#endif
    }
    if (own_terms) be_free(dterms);
    if (rc != 0) return fail(LITHO_ERR_CUDA, std::string("pupil_build: ") + be_errstr(rc));
    return LITHO_OK;
}

// FP32 FMA-throughput probe used by bench.py for the roofline denominator: every thread runs
// `iters` rounds of 16 independent FMAs.  Returns the flop count of the launch in *flops.
int litho_fp32_probe(float* out, int blocks, int iters, double* flops, void* stream) {
    if (!out || blocks <= 0 || iters <= 0) return fail(LITHO_ERR_ARG, "fp32_probe: bad argument");
    if (flops) *flops = (double)blocks * 256.0 * (double)iters * 16.0 * 2.0;
#if defined(LITHO_EMU)
    (void)stream;
    out[0] = 0.f;
#else
    fma_probe_kernel<<<blocks, 256, 0, (litho_stream_t)stream>>>(out, iters);
    BE_CHECK((int)cudaGetLastError());
#endif
    return LITHO_OK;
}

}  // extern "C"
