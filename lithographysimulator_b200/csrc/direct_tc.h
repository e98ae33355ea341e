// Tensor-core (tcgen05, 3xTF32) form of the direct solver's two complex matrix products; see direct_tc.cu.
// Device build only: the CPU emulation of the kernels (tests/emu) keeps the FP32 CUDA-core path.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

namespace litho_tc {

typedef float2 cplx2;

struct TcParams {
    const cplx2* A;        // [pn][pn] operator, row-major
    int pn;
    const cplx2* pupil;    // stage 1: [pn][pn]
    const cplx2* mask;     // stage 1: [pn][pn]
    int pr0, pc0, Sr, Sc;  // pupil window
    const int2* shifts;    // source shifts (may be null: zero shift)
    int s_begin;
    int stage;             // 1: U_s[b][u] = sum_v A[b][col(v)] G_s[u][v];  2: |sum_u A[a][row(u)] U_s[b][u]|^2
    int NT;                // complex columns per N tile (set by tc_launch)
    const cplx2* U;        // stage 2 input  [batch][pn][Upitch]
    cplx2* Uout;           // stage 1 output [batch][pn][Upitch]
    int Upitch;
    float* part;           // stage 2 output [batch][pn][pn]: |E_s|^2 of each source point
    int* err;              // device word set to 1 if an mbarrier wait timed out (results invalid)
};

int tc_ctas_per_point(int pn, int n_out);   // CTAs one source point occupies in the stage with n_out output columns
int tc_launch(TcParams P, int n_out, int batch, cudaStream_t st);
int tc_reduce(const float* part, const float* weights, int s_begin, int batch, int pn, float* intensity, cudaStream_t st);

}  // namespace litho_tc
