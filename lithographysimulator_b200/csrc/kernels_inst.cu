// Instantiation unit: compiled once per sub-FFT length with -DLITHO_INST_M=<M>.
#include "launch.h"

#ifndef LITHO_INST_M
#error "compile with -DLITHO_INST_M=<sub-FFT length>"
#endif

#if defined(LITHO_EMU)
#include "emu_runtime.h"
#endif

namespace litho {

#if !defined(LITHO_EMU)
template <int M, int KIND>
__global__ void __launch_bounds__(RowsShape<M>::THREADS) abbe_rows_kernel(const __grid_constant__ RowsParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    rows_body<M, KIND>(P, DevCtx{}, reinterpret_cast<cplx*>(smem_raw));
}

template <int M, int EPI>
__global__ void __launch_bounds__(ColsShape<M>::THREADS) abbe_cols_kernel(const __grid_constant__ ColsParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cols_body<M, EPI>(P, DevCtx{}, reinterpret_cast<cplx*>(smem_raw));
}

template <class K>
static int set_smem(K kernel, size_t bytes) {
    if (bytes > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        if (e != cudaSuccess) return (int)e;
    }
    return 0;
}
#endif

template <int M>
int launch_rows_m(int kind, const RowsParams& P, int gx, int gy, litho_stream_t st) {
    const size_t smem = (size_t)RowsShape<M>::FPC * FftShape<M>::SMEM_ELEMS * sizeof(cplx);
    const int threads = RowsShape<M>::THREADS;
#if defined(LITHO_EMU)
    (void)st;
    auto run = [&](auto body) { litho_emu::launch(gx, gy, 1, threads, smem, body); };
    if (kind == ROW_PUPIL_MASK)
        run([&](const litho_emu::EmuCtx& c, char* s) { rows_body<M, ROW_PUPIL_MASK>(P, c, (cplx*)s); });
    else if (kind == ROW_REAL_PLANE)
        run([&](const litho_emu::EmuCtx& c, char* s) { rows_body<M, ROW_REAL_PLANE>(P, c, (cplx*)s); });
    else
        run([&](const litho_emu::EmuCtx& c, char* s) { rows_body<M, ROW_CPLX_PLANE>(P, c, (cplx*)s); });
    return 0;
#else
    dim3 grid(gx, gy, 1), block(threads, 1, 1);
    int e = 0;
    if (kind == ROW_PUPIL_MASK) {
        if ((e = set_smem(abbe_rows_kernel<M, ROW_PUPIL_MASK>, smem))) return e;
        abbe_rows_kernel<M, ROW_PUPIL_MASK><<<grid, block, smem, st>>>(P);
    } else if (kind == ROW_REAL_PLANE) {
        if ((e = set_smem(abbe_rows_kernel<M, ROW_REAL_PLANE>, smem))) return e;
        abbe_rows_kernel<M, ROW_REAL_PLANE><<<grid, block, smem, st>>>(P);
    } else {
        if ((e = set_smem(abbe_rows_kernel<M, ROW_CPLX_PLANE>, smem))) return e;
        abbe_rows_kernel<M, ROW_CPLX_PLANE><<<grid, block, smem, st>>>(P);
    }
    return (int)cudaGetLastError();
#endif
}

template <int M>
int launch_cols_m(int epi, const ColsParams& P, int gx, int gy, litho_stream_t st) {
    const size_t smem = (size_t)ColsShape<M>::CB * FftShape<M>::SMEM_ELEMS * sizeof(cplx);
    const int threads = ColsShape<M>::THREADS;
#if defined(LITHO_EMU)
    (void)st;
    auto run = [&](auto body) { litho_emu::launch(gx, gy, 1, threads, smem, body); };
    if (epi == EPI_ACCUM)
        run([&](const litho_emu::EmuCtx& c, char* s) { cols_body<M, EPI_ACCUM>(P, c, (cplx*)s); });
    else
        run([&](const litho_emu::EmuCtx& c, char* s) { cols_body<M, EPI_FIELD>(P, c, (cplx*)s); });
    return 0;
#else
    dim3 grid(gx, gy, 1), block(threads, 1, 1);
    int e = 0;
    if (epi == EPI_ACCUM) {
        if ((e = set_smem(abbe_cols_kernel<M, EPI_ACCUM>, smem))) return e;
        abbe_cols_kernel<M, EPI_ACCUM><<<grid, block, smem, st>>>(P);
    } else {
        if ((e = set_smem(abbe_cols_kernel<M, EPI_FIELD>, smem))) return e;
        abbe_cols_kernel<M, EPI_FIELD><<<grid, block, smem, st>>>(P);
    }
    return (int)cudaGetLastError();
#endif
}

template <int M>
void shape_m(int* rows_fpc, int* cols_cb) {
    *rows_fpc = RowsShape<M>::FPC;
    *cols_cb = ColsShape<M>::CB;
}

#if LITHO_INST_M >= 32 && LITHO_INST_M <= 4096
#if !defined(LITHO_EMU)
template <int M, int PPT>
__global__ void __launch_bounds__(FastShape<M, PPT>::ROW_THREADS, FastShape<M, PPT>::ROW_MIN_BLOCKS)
abbe_fast_rows_kernel(const __grid_constant__ FastRowsParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    fast_rows_body<M, PPT>(P, DevCtx{}, reinterpret_cast<cplx*>(smem_raw));
}
template <int M, int PPT>
__global__ void __launch_bounds__(FastShape<M, PPT>::COL_THREADS, FastShape<M, PPT>::COL_MIN_BLOCKS)
abbe_fast_cols_kernel(const __grid_constant__ FastColsParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    fast_cols_body<M, PPT>(P, DevCtx{}, reinterpret_cast<cplx*>(smem_raw));
}
template <int M, int PPT, int CBT>
__global__ void __launch_bounds__(TmaShape<M, PPT, CBT>::THREADS, TmaShape<M, PPT, CBT>::MIN_BLOCKS)
abbe_fast_cols_tma_kernel(const __grid_constant__ FastColsParams P) {
    extern __shared__ __align__(128) unsigned char smem_raw_tma[];
    fast_cols_tma_body<M, PPT, CBT>(P, DevCtx{}, smem_raw_tma);
}
#endif

// wide / narrow tile widths of the TMA-staged column kernel for this M (0: none)
template <int M, int PPT>
struct TmaWidths {
    using F = FastShape<M, PPT>;
    static constexpr int WIDE_C = F::CB_DUAL >= 2 ? F::CB_DUAL : 2;
    static constexpr int WIDE = (F::CB_DUAL >= 2 && TmaShape<M, PPT, WIDE_C>::OK) ? WIDE_C : 0;
    // the narrow tile only pays when two CTAs then fit an SM
    static constexpr int NARROW_C = F::CB_DUAL / 2 >= 2 ? F::CB_DUAL / 2 : 2;
    static constexpr int NARROW = (F::CB_DUAL / 2 >= 2 && TmaShape<M, PPT, NARROW_C>::OK &&
                                   TmaShape<M, PPT, NARROW_C>::MIN_BLOCKS == 2) ? NARROW_C : 0;
};

template <int M, int PPT, int CBT>
int launch_fast_cols_tma(const FastColsParams& P, litho_stream_t st) {
    using S = TmaShape<M, PPT, CBT>;
    const int gx = 2 * (M / CBT);
#if defined(LITHO_EMU)
    (void)st;
    litho_emu::launch(gx, 1, 1, S::THREADS, S::SMEM, [&](const litho_emu::EmuCtx& c, char* s) {
        fast_cols_tma_body<M, PPT, CBT>(P, c, (unsigned char*)s);
    });
    return 0;
#else
    int e = set_smem(abbe_fast_cols_tma_kernel<M, PPT, CBT>, S::SMEM);
    if (e) return e;
    abbe_fast_cols_tma_kernel<M, PPT, CBT><<<dim3(gx, 1, 1), dim3(S::THREADS, 1, 1), S::SMEM, st>>>(P);
    return (int)cudaGetLastError();
#endif
}

template <int M, int PPT>
int launch_fast_rows_m(const FastRowsParams& P, int gx, litho_stream_t st) {
    using F = FastShape<M, PPT>;
#if defined(LITHO_EMU)
    (void)st;
    litho_emu::launch(gx, 1, 1, F::ROW_THREADS, F::ROW_SMEM,
                      [&](const litho_emu::EmuCtx& c, char* s) { fast_rows_body<M, PPT>(P, c, (cplx*)s); });
    return 0;
#else
    int e = set_smem(abbe_fast_rows_kernel<M, PPT>, F::ROW_SMEM);
    if (e) return e;
    abbe_fast_rows_kernel<M, PPT><<<dim3(gx, 1, 1), dim3(F::ROW_THREADS, 1, 1), F::ROW_SMEM, st>>>(P);
    return (int)cudaGetLastError();
#endif
}

template <int M, int PPT>
int launch_fast_cols_m(const FastColsParams& P, litho_stream_t st) {
    using F = FastShape<M, PPT>;
    const int gx = 2 * (M / F::CB);
    if (P.use_tma) {  // P.use_tma = columns per tile
        using W = TmaWidths<M, PPT>;
        if constexpr (W::WIDE > 0) {
            if (P.use_tma == W::WIDE) return launch_fast_cols_tma<M, PPT, W::WIDE>(P, st);
        }
        if constexpr (W::NARROW > 0) {
            if (P.use_tma == W::NARROW) return launch_fast_cols_tma<M, PPT, W::NARROW>(P, st);
        }
        return -2;
    }
#if defined(LITHO_EMU)
    (void)st;
    litho_emu::launch(gx, F::COL_GRID_Y, 1, F::COL_THREADS, F::COL_SMEM,
                      [&](const litho_emu::EmuCtx& c, char* s) { fast_cols_body<M, PPT>(P, c, (cplx*)s); });
    return 0;
#else
    int e = set_smem(abbe_fast_cols_kernel<M, PPT>, F::COL_SMEM);
    if (e) return e;
    abbe_fast_cols_kernel<M, PPT><<<dim3(gx, F::COL_GRID_Y, 1), dim3(F::COL_THREADS, 1, 1), F::COL_SMEM, st>>>(P);
    return (int)cudaGetLastError();
#endif
}

template <int M, int PPT>
int fast_ntab_m() {
    return FastShape<M, PPT>::NTAB;
}
// TMA-staged column kernel of this M: which = 0 wide tile width, 1 narrow tile width (0: not available),
// 2 rows per TMA box, 3 element count of the compact twiddle table
template <int M, int PPT>
int fast_tma_cols_m(int which) {
    using W = TmaWidths<M, PPT>;
    using S = TmaShape<M, PPT, (W::WIDE > 0 ? W::WIDE : 2)>;
    switch (which) {
        case 0: return W::WIDE;
        case 1: return W::NARROW;
        case 2: return S::BOX_ROWS;
        default: return S::NTAB_PAD;
    }
}

template int launch_fast_rows_m<LITHO_INST_M, 32>(const FastRowsParams&, int, litho_stream_t);
template int launch_fast_cols_m<LITHO_INST_M, 32>(const FastColsParams&, litho_stream_t);
template int fast_ntab_m<LITHO_INST_M, 32>();
template int fast_tma_cols_m<LITHO_INST_M, 32>(int);
#endif

template int launch_rows_m<LITHO_INST_M>(int, const RowsParams&, int, int, litho_stream_t);
template int launch_cols_m<LITHO_INST_M>(int, const ColsParams&, int, int, litho_stream_t);
template void shape_m<LITHO_INST_M>(int*, int*);

}  // namespace litho
