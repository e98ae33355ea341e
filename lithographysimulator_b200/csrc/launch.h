// Launch layer: one translation unit per sub-FFT length M (kernels_inst.cu is compiled once
// per M with -DLITHO_INST_M=<M>) so the heavily unrolled kernels build in parallel.
// With -DLITHO_EMU the same entry points run the kernel bodies on the CPU (tests/emu).
#pragma once
#include "abbe_kernels.h"

#if defined(LITHO_EMU)
typedef void* litho_stream_t;
#else
#include <cuda_runtime.h>
typedef cudaStream_t litho_stream_t;
#endif

namespace litho {

// return 0 on success, otherwise a cudaError_t value (device build)
template <int M>
int launch_rows_m(int kind, const RowsParams& P, int gx, int gy, litho_stream_t st);
template <int M>
int launch_cols_m(int epi, const ColsParams& P, int gx, int gy, litho_stream_t st);
template <int M>
void shape_m(int* rows_fpc, int* cols_cb);

#define LITHO_FOR_EACH_M(X) X(16) X(32) X(64) X(128) X(256) X(512) X(1024) X(2048) X(4096) X(8192) X(16384)

}  // namespace litho

// ---- fast path (fast_kernels.h), instantiated for 32 <= M <= 4096 ----
#include "fast_kernels.h"
namespace litho {
template <int M, int PPT>
int launch_fast_rows_m(const FastRowsParams& P, int gx, litho_stream_t st);
template <int M, int PPT>
int launch_fast_cols_m(const FastColsParams& P, litho_stream_t st);
template <int M, int PPT>
int fast_ntab_m();
template <int M, int PPT>
int fast_tma_cols_m(int which);
#define LITHO_FOR_EACH_FAST_M(X) X(32) X(64) X(128) X(256) X(512) X(1024) X(2048) X(4096)
}  // namespace litho
