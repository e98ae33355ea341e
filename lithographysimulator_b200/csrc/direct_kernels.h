// Direct ("Abbe") solver -- reference imageformation.py:3-30 and mask.py:41-61.
//
// The reference evaluates a double trapezoid-rule Fourier integral through a [pn,pn,pn,pn]
// tensor.  The integrand separates:
//     E = A * G * A^T,   A[a][c] = w[c] * exp(sign * i * (2*pi/lambda) * Q[a][c]),
//     Q[a][c] = fp16( fp16(k[a]) * fp16(x[c]) ),  w = trapezoid weights (1/2 at both ends),
// so one source point is two complex matrix products restricted to the non-zero window of
// G_s = roll(P, shift_s) * M.  The fp16-quantised phase table must be reproduced exactly; exact
// phases are 1e-3 away from the reference.
//
//   direct_op_body   : builds A (once per call)
//   direct_rows_body : T_s[u][b] = sum_v G_s[u][v] * A[b][col(v)]          (window rows u)
//   direct_cols_body : E_s[a][b] = sum_u A[a][row(u)] * T_s[u][b], then |E|^2 summed over the
//                      batch in registers and added to the intensity (or E stored as a field).
#pragma once
#include "hd.h"
#if defined(__CUDACC__)
#include <cuda_fp16.h>
#endif

namespace litho {

LITHO_HD float round_f16(float v) {
#if defined(__CUDA_ARCH__)
    return __half2float(__float2half_rn(v));
#else
    // round-to-nearest-even float -> half -> float on the host (normal, subnormal, overflow)
    union { float f; uint32_t u; } in, out;
    in.f = v;
    const uint32_t sign = in.u & 0x80000000u;
    uint32_t ax = in.u & 0x7fffffffu;
    if (ax >= 0x7f800000u) return v;                  // inf / nan
    if (ax >= 0x477ff000u) {                          // rounds to >= 65520 -> inf
        out.u = sign | 0x7f800000u;
        return out.f;
    }
    if (ax < 0x38800000u) {                           // half subnormal range: quantum 2^-24
        const float q = 5.9604644775390625e-08f;
        float a = in.f < 0 ? -in.f : in.f;
        float r = (float)nearbyint((double)a / (double)q) * q;
        return in.f < 0 ? -r : r;
    }
    const uint32_t lsb = (ax >> 13) & 1u;
    ax += 0x00000fffu + lsb;
    ax &= 0xffffe000u;
    out.u = sign | ax;
    return out.f;
#endif
}

struct DirectOpParams {
    float kstart, kstep, xstart, xstep;  // float32(start), float32(step) of the two fp16 aranges
    float c0;                            // float32(2*pi/lambda)
    int sign;                            // -1: imaging (imageformation.py:52), +1: mask spectrum (mask.py:42)
    int pn;
    cplx* A;
};

LITHO_HD void direct_op_elem(const DirectOpParams& P, int a, int c) {
    const float k = round_f16(P.kstart + (float)a * P.kstep);
    const float x = round_f16(P.xstart + (float)c * P.xstep);
    const float q = round_f16(k * x);
    const double ang = (double)P.c0 * (double)q;
    const double w = (c == 0 || c == P.pn - 1) ? 0.5 : 1.0;
    P.A[(size_t)a * P.pn + c] = mk((float)(w * cos(ang)), (float)(w * (double)P.sign * sin(ang)));
}

enum DirectKind { DIRECT_PUPIL_MASK = 0, DIRECT_GEOMETRY = 1 };
enum DirectEpi { DIRECT_ACCUM = 0, DIRECT_FIELD = 1 };

struct DirectParams {
    const cplx* A;
    int pn;
    const cplx* pupil;
    const cplx* mask;
    const int16_t* geometry;
    int pr0, pc0, Sr, Sc;  // non-zero window of the pupil (full grid for DIRECT_GEOMETRY)
    const int2_* shifts;
    const float* weights;
    int s_begin, batch;
    cplx* T;  // [batch][Sr][pn]
    float* intensity;  // [pn][pn], accumulated
    cplx* field;       // [pn][pn]
};

constexpr int DT = 32;   // output tile side
constexpr int DK = 16;   // k-chunk
constexpr int DIRECT_THREADS = 256;
constexpr int DIRECT_SMEM_ELEMS = 2 * DT * (DK + 1);

// T_s[u][b] = sum_v G_s[u][v] * A[b][col(v)]      grid: (ceil(pn/32), ceil(Sr/32), batch)
template <int KIND, class Ctx>
LITHO_HD void direct_rows_body(const DirectParams& P, const Ctx& ctx, cplx* smem) {
    cplx (*Gs)[DK + 1] = reinterpret_cast<cplx (*)[DK + 1]>(smem);
    cplx (*As)[DK + 1] = reinterpret_cast<cplx (*)[DK + 1]>(smem + DT * (DK + 1));
    const int tid = ctx.tid();
    const int tx = tid & 15, ty = tid >> 4;
    const int b0 = ctx.bx() * DT, u0 = ctx.by() * DT, sl = ctx.bz();
    int d0 = 0, d1 = 0;
    if (P.shifts) {
        int2_ sh = P.shifts[P.s_begin + sl];
        d0 = sh.x; d1 = sh.y;
    }
    const int pn = P.pn;
    cplx acc[2][2];
    acc[0][0] = acc[0][1] = acc[1][0] = acc[1][1] = mk(0.f, 0.f);
    for (int v0 = 0; v0 < P.Sc; v0 += DK) {
        ctx.sync();
        // stage a 32 x 16 block of G and of A: 512 elements each, 2 per thread
        for (int i = tid; i < DT * DK; i += DIRECT_THREADS) {
            const int r = i / DK, kx = i - r * DK;
            const int u = u0 + r, v = v0 + kx, b = b0 + r;
            cplx g = mk(0.f, 0.f), a = mk(0.f, 0.f);
            if (v < P.Sc) {
                const int col = imod(P.pc0 + d1 + v, pn);  // true grid column of window column v
                if (u < P.Sr) {
                    if (KIND == DIRECT_PUPIL_MASK) {
                        const int row = imod(P.pr0 + d0 + u, pn);
                        g = cmul(P.pupil[(size_t)(P.pr0 + u) * pn + P.pc0 + v], P.mask[(size_t)row * pn + col]);
                    } else {
                        g = mk((float)P.geometry[(size_t)(P.pr0 + u) * pn + P.pc0 + v], 0.f);
                    }
                }
                if (b < pn) a = P.A[(size_t)b * pn + col];
            }
            Gs[r][kx] = g;
            As[r][kx] = a;
        }
        ctx.sync();
#pragma unroll
        for (int kx = 0; kx < DK; ++kx) {
            const cplx g0 = Gs[ty][kx], g1 = Gs[ty + 16][kx];
            const cplx a0 = As[tx][kx], a1 = As[tx + 16][kx];
            acc[0][0] = cadd(acc[0][0], cmul(g0, a0));
            acc[0][1] = cadd(acc[0][1], cmul(g0, a1));
            acc[1][0] = cadd(acc[1][0], cmul(g1, a0));
            acc[1][1] = cadd(acc[1][1], cmul(g1, a1));
        }
    }
    cplx* T = P.T + (size_t)sl * P.Sr * pn;
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int u = u0 + ty + 16 * i, b = b0 + tx + 16 * j;
            if (u < P.Sr && b < pn) T[(size_t)u * pn + b] = acc[i][j];
        }
}

// E_s[a][b] = sum_u A[a][row(u)] * T_s[u][b]       grid: (ceil(pn/32), ceil(pn/32), 1)
template <int EPI, class Ctx>
LITHO_HD void direct_cols_body(const DirectParams& P, const Ctx& ctx, cplx* smem) {
    cplx (*As)[DK + 1] = reinterpret_cast<cplx (*)[DK + 1]>(smem);            // [a][k]
    cplx (*Ts)[DT + 1] = reinterpret_cast<cplx (*)[DT + 1]>(smem + DT * (DK + 1));  // [k][b]
    const int tid = ctx.tid();
    const int tx = tid & 15, ty = tid >> 4;
    const int b0 = ctx.bx() * DT, a0 = ctx.by() * DT;
    const int pn = P.pn;
    float inten[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
    cplx acc[2][2];
    for (int sl = 0; sl < P.batch; ++sl) {
        int d0 = 0;
        if (P.shifts) d0 = P.shifts[P.s_begin + sl].x;
        const cplx* T = P.T + (size_t)sl * P.Sr * pn;
        acc[0][0] = acc[0][1] = acc[1][0] = acc[1][1] = mk(0.f, 0.f);
        for (int k0 = 0; k0 < P.Sr; k0 += DK) {
            ctx.sync();
            for (int i = tid; i < DT * DK; i += DIRECT_THREADS) {
                {   // A block: 32 (a) x 16 (k)
                    const int r = i / DK, kx = i - r * DK;
                    const int a = a0 + r, u = k0 + kx;
                    cplx v = mk(0.f, 0.f);
                    if (a < pn && u < P.Sr) v = P.A[(size_t)a * pn + imod(P.pr0 + d0 + u, pn)];
                    As[r][kx] = v;
                }
                {   // T block: 16 (k) x 32 (b)
                    const int kx = i / DT, c = i - kx * DT;
                    const int u = k0 + kx, b = b0 + c;
                    cplx v = mk(0.f, 0.f);
                    if (u < P.Sr && b < pn) v = T[(size_t)u * pn + b];
                    Ts[kx][c] = v;
                }
            }
            ctx.sync();
#pragma unroll
            for (int kx = 0; kx < DK; ++kx) {
                const cplx x0 = As[ty][kx], x1 = As[ty + 16][kx];
                const cplx y0 = Ts[kx][tx], y1 = Ts[kx][tx + 16];
                acc[0][0] = cadd(acc[0][0], cmul(x0, y0));
                acc[0][1] = cadd(acc[0][1], cmul(x0, y1));
                acc[1][0] = cadd(acc[1][0], cmul(x1, y0));
                acc[1][1] = cadd(acc[1][1], cmul(x1, y1));
            }
        }
        if (EPI == DIRECT_ACCUM) {
            const float w = P.weights ? P.weights[P.s_begin + sl] : 1.f;
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int j = 0; j < 2; ++j) inten[i][j] += w * cnorm2(acc[i][j]);
        }
    }
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int a = a0 + ty + 16 * i, b = b0 + tx + 16 * j;
            if (a < pn && b < pn) {
                if (EPI == DIRECT_ACCUM) P.intensity[(size_t)a * pn + b] += inten[i][j];
                else P.field[(size_t)a * pn + b] = acc[i][j];
            }
        }
}

}  // namespace litho
