// Kernel bodies of the Abbe FFT-approximation path (reference imageformation.py:32-77).
//
// Two passes per source point, both built on fft_core.h / zoom.h:
//   rows_body : for every non-zero row of G_s = roll(P, shift_s) * M (formed on the fly from
//               the pupil and mask-spectrum planes, never stored) the pruned zoom DFT along the
//               row -> T_s (residue-major, only the wanted outputs).
//   cols_body : for every wanted output column the pruned zoom DFT along the column of T_s,
//               then |E|^2, weighted, summed over the batch of source points in registers and
//               added once to the residue-major intensity plane (or stored as a complex field).
// The [n_source, N, N] stack of the naive formulation is never materialised: the only
// intermediate is T (S x W complex per in-flight source point).
//
// Bodies are templated on a thread context so tests/emu can run them on the CPU (hd.h).
#pragma once
#include "fft_core.h"
#include "zoom.h"

namespace litho {

enum RowKind { ROW_PUPIL_MASK = 0, ROW_REAL_PLANE = 1, ROW_CPLX_PLANE = 2 };
enum ColEpi { EPI_ACCUM = 0, EPI_FIELD = 1 };

struct RowsParams {
    // ROW_PUPIL_MASK: G[line][u] = pupil[pr0+line][pc0+u] * mask[(pr0+line+d0) mod pn][(pc0+u+d1) mod pn]
    const cplx* pupil;
    const cplx* mask;
    int pn;
    int pr0, pc0;
    // ROW_REAL_PLANE / ROW_CPLX_PLANE: in[line*in_pitch + u]
    const float* real_in;
    const cplx* cplx_in;
    int in_pitch;
    // per-source shifts (d0,d1) = argwhere(lightsource) - pn//2  (imageformation.py:59); may be null
    const int2_* shifts;
    int s_begin;
    int lines;    // number of lines transformed per source point
    AxisIn ax;    // along-line input axis (first excludes the per-source shift)
    AxisOut out;  // wanted outputs along the line
    ZoomPlan plan;
    const cplx* twL;
    cplx* T;  // [batch][R][lines][Wr]
};

struct ColsParams {
    const cplx* T;  // [batch][Rc][S][Wrc]   (S = ax.S lines, column residues of the row pass)
    int batch;
    int Rc, Wrc;     // residue count / pitch of the row pass (column index space)
    AxisOut outc;    // wanted outputs of the row pass (gives cnt per column residue)
    const int2_* shifts;
    const float* weights;  // per-source weight or null (reference: all ones, Q2)
    int s_begin;
    AxisIn ax;    // along-column input axis (first excludes the per-source shift)
    AxisOut out;  // wanted outputs along the column
    ZoomPlan plan;
    const cplx* twL;
    // EPI_ACCUM: iperm[((rr*Rc + rc)*Wr + kkr)*Wrc + kkc] += sum_s w_s |E|^2
    float* iperm;
    // EPI_FIELD: field[o_r*field_pitch + o_c] = E (conjugated if conj_out)
    cplx* field;
    float* real_out;  // when non-null only the real part is stored here (same pitch)
    int field_pitch;
    int conj_out;
    float scale;
};

template <int M>
struct ColsShape {
    static constexpr int TG = FftShape<M>::TG;
    static constexpr int CB = (512 / TG) >= 8 ? 8 : ((512 / TG) >= 1 ? (512 / TG) : 1);  // columns per CTA
    static constexpr int THREADS = CB * TG;
};

template <int M>
struct RowsShape {
    static constexpr int TG = FftShape<M>::TG;
    static constexpr int FPC = (256 / TG) >= 1 ? (256 / TG) : 1;  // FFT groups (work items) per CTA
    static constexpr int THREADS = FPC * TG;
};

// ----------------------------------------------------------------------------- rows pass
// grid.x = ceil(lines*R / FPC), grid.y = batch.   work item = line*R + r.
template <int M, int KIND, class Ctx>
LITHO_HD void rows_body(const RowsParams& P, const Ctx& ctx, cplx* smem) {
    using Sh = FftShape<M>;
    constexpr int TG = Sh::TG;
    constexpr int FPC = RowsShape<M>::FPC;
    const int grp = ctx.tid() / TG;
    const int g = ctx.tid() - grp * TG;
    const int sl = ctx.by();
    const int R = P.plan.R;
    const int item = ctx.bx() * FPC + grp;
    const bool active = item < P.lines * R;
    const int line = active ? item / R : 0;
    const int r = active ? item - line * R : 0;

    int d0 = 0, d1 = 0;
    if (P.shifts) {
        int2_ sh = P.shifts[P.s_begin + sl];
        d0 = sh.x;
        d1 = sh.y;
    }
    AxisIn ax = P.ax;
    ax.first += d1;

    cplx v[16];
    if (active) {
        if constexpr (KIND == ROW_PUPIL_MASK) {
            const cplx* prow = P.pupil + (size_t)(P.pr0 + line) * P.pn + P.pc0;
            const cplx* mrow = P.mask + (size_t)imod(P.pr0 + line + d0, P.pn) * P.pn;
            const int mc0 = imod(P.pc0 + d1, P.pn);
            const int pn = P.pn;
            auto ld = [&](int u) {
                int mc = mc0 + u;
                if (mc >= pn) mc -= pn;
                return cmul(ldg_c(prow + u), ldg_c(mrow + mc));
            };
            zoom_load<16>(v, g, TG, M, P.plan.L, r, ax, P.twL, ld);
        } else if constexpr (KIND == ROW_REAL_PLANE) {
            const float* row = P.real_in + (size_t)line * P.in_pitch;
            auto ld = [&](int u) { return mk(ldg_f(row + u), 0.f); };
            zoom_load<16>(v, g, TG, M, P.plan.L, r, ax, P.twL, ld);
        } else {
            const cplx* row = P.cplx_in + (size_t)line * P.in_pitch;
            auto ld = [&](int u) { return ldg_c(row + u); };
            zoom_load<16>(v, g, TG, M, P.plan.L, r, ax, P.twL, ld);
        }
    } else {
#pragma unroll
        for (int e = 0; e < 16; ++e) v[e] = mk(0.f, 0.f);
    }

    fft_run<M, 16, false>(v, smem + grp * Sh::SMEM_ELEMS, 1, g, GlobalTw<M, 16>{P.twL, R}, CtaSync<Ctx>{ctx});

    if (active) {
        const int kmin = zoom_kmin(P.out, R, r);
        const int cnt = zoom_kend(P.out, R, r) - kmin;
        cplx* dst = P.T + ((size_t)(sl * R + r) * P.lines + line) * P.plan.Wr;
#pragma unroll
        for (int e = 0; e < 16; ++e) {
            const int kk = (g + TG * e - kmin) & (M - 1);
            if (kk < cnt) dst[kk] = v[e];
        }
    }
}

// ----------------------------------------------------------------------------- cols pass
// grid.x = Rc * ceil(Wrc / CB) column blocks, grid.y = R (row residue).  thread = col + CB*g.
template <int M, int EPI, class Ctx>
LITHO_HD void cols_body(const ColsParams& P, const Ctx& ctx, cplx* smem) {
    using Sh = FftShape<M>;
    constexpr int TG = Sh::TG;
    constexpr int CB = ColsShape<M>::CB;
    const int col = ctx.tid() % CB;
    const int g = ctx.tid() / CB;
    const int R = P.plan.R;
    const int rr = ctx.by();
    const int nblk = (P.Wrc + CB - 1) / CB;
    const int rc = ctx.bx() / nblk;
    const int kkc = (ctx.bx() - rc * nblk) * CB + col;
    const int kminc = zoom_kmin(P.outc, P.Rc, rc);
    const bool colvalid = kkc < zoom_kend(P.outc, P.Rc, rc) - kminc;
    const int S = P.ax.S;

    float acc[16];
#pragma unroll
    for (int e = 0; e < 16; ++e) acc[e] = 0.f;
    cplx v[16];

    for (int sl = 0; sl < P.batch; ++sl) {
        AxisIn ax = P.ax;
        if (P.shifts) ax.first += P.shifts[P.s_begin + sl].x;
        if (colvalid) {
            const cplx* src = P.T + ((size_t)(sl * P.Rc + rc) * S) * P.Wrc + kkc;
            const int pitch = P.Wrc;
            auto ld = [&](int u) { return ldg_c(src + (size_t)u * pitch); };
            zoom_load<16>(v, g, TG, M, P.plan.L, rr, ax, P.twL, ld);
        } else {
#pragma unroll
            for (int e = 0; e < 16; ++e) v[e] = mk(0.f, 0.f);
        }

        fft_run<M, 16, false>(v, smem + col, CB, g, GlobalTw<M, 16>{P.twL, R}, CtaSync<Ctx>{ctx});

        if constexpr (EPI == EPI_ACCUM) {
            const float w = P.weights ? P.weights[P.s_begin + sl] : 1.f;
#pragma unroll
            for (int e = 0; e < 16; ++e) acc[e] += w * cnorm2(v[e]);
        }
    }

    if (!colvalid) return;
    const int kmin = zoom_kmin(P.out, R, rr);
    const int cnt = zoom_kend(P.out, R, rr) - kmin;
    if constexpr (EPI == EPI_ACCUM) {
        float* dst = P.iperm + ((size_t)(rr * P.Rc + rc) * P.plan.Wr) * P.Wrc + kkc;
#pragma unroll
        for (int e = 0; e < 16; ++e) {
            const int kk = (g + TG * e - kmin) & (M - 1);
            if (kk < cnt) dst[(size_t)kk * P.Wrc] += acc[e];
        }
    } else {
        const int oc = P.Rc * (kminc + kkc) + rc + P.outc.center;
#pragma unroll
        for (int e = 0; e < 16; ++e) {
            const int kk = (g + TG * e - kmin) & (M - 1);
            if (kk < cnt) {
                const int orow = R * (kmin + kk) + rr + P.out.center;
                cplx o = v[e];
                o.x *= P.scale;
                o.y *= P.conj_out ? -P.scale : P.scale;
                if (P.real_out) P.real_out[(size_t)orow * P.field_pitch + oc] = o.x;
                else P.field[(size_t)orow * P.field_pitch + oc] = o;
            }
        }
    }
}

// ----------------------------------------------------------------------------- finalize
// Natural-order read of the residue-major intensity plane.
struct PermView {
    int mode;  // 0: residue-major plane of the generic path, 1: natural row-major plane,
               // 2: coarse plane of the fast path [2][2][M][M] when Nc == N (no interpolation needed)
    const float* iperm;
    AxisOut outr, outc;  // row / column output axes (W = pn, center = pn//2)
    int Rr, Rc, Wrr, Wrc;
    int pitch;           // mode 1
    int Mc, center;      // mode 2: sub-FFT length and pn//2; fine index i' = i - center, o = i' mod 2M
    LITHO_HD float at(int i, int j) const {
        if (mode == 1) return iperm[(size_t)i * pitch + j];
        if (mode == 2) {
            const int oi = (i - center) & (2 * Mc - 1), oj = (j - center) & (2 * Mc - 1);
            return iperm[((size_t)((oi & 1) * 2 + (oj & 1)) * Mc + (oi >> 1)) * Mc + (oj >> 1)];
        }
        int rr, kr, rc, kc;
        zoom_split(outr, Rr, i, rr, kr);
        zoom_split(outc, Rc, j, rc, kc);
        return iperm[((size_t)(rr * Rc + rc) * Wrr + kr) * Wrc + kc];
    }
};

// imageformation.py:69-75: abs, bilinear resample by 1/eps (align_corners=False, given scale
// factor), zero border of (pW, pW+corr).  One thread per output pixel; `scale` is float32(eps)
// and the source index uses a single-rounding FMA like ATen.
struct FinalizeParams {
    PermView in;
    int pn;         // input side
    int side;       // resampled side = floor(pn * (1/eps))
    int pW;         // leading zero border
    int out_side;   // side + 2*pW + corr
    float scale;    // float32(1 / (1/eps))
    float* out;     // [out_side][out_side]
};

LITHO_HD void bilinear_src(float scale, int dst, int in_size, int& i0, int& i1, float& l0, float& l1) {
#if defined(__CUDA_ARCH__)
    float src = fmaf(scale, (float)dst + 0.5f, -0.5f);
#else
    float src = (float)((double)scale * ((double)dst + 0.5) - 0.5);
#endif
    if (src < 0.f) src = 0.f;
    i0 = (int)src;
    if (i0 > in_size - 1) i0 = in_size - 1;
    i1 = i0 + 1 < in_size ? i0 + 1 : in_size - 1;
    l1 = src - (float)i0;
    l0 = 1.f - l1;
}

// mask.py:76-77: int16 geometry -> float32, bilinear resample by eps (same ATen rules as above)
struct ResampleParams {
    const int16_t* in;  // [pn][pn]
    int pn;
    int side;     // floor(pn * eps)
    float scale;  // float32(1 / eps)
    float* out;   // [side][side]
};

LITHO_HD void resample_pixel(const ResampleParams& P, int y, int x) {
    float val;
    if (P.side == P.pn) {
        val = (float)P.in[(size_t)y * P.pn + x];
    } else {
        int r0, r1, c0, c1;
        float lh0, lh1, lw0, lw1;
        bilinear_src(P.scale, y, P.pn, r0, r1, lh0, lh1);
        bilinear_src(P.scale, x, P.pn, c0, c1, lw0, lw1);
        const float a = (float)P.in[(size_t)r0 * P.pn + c0], b = (float)P.in[(size_t)r0 * P.pn + c1];
        const float c = (float)P.in[(size_t)r1 * P.pn + c0], d = (float)P.in[(size_t)r1 * P.pn + c1];
        val = lh0 * (lw0 * a + lw1 * b) + lh1 * (lw0 * c + lw1 * d);
    }
    P.out[(size_t)y * P.side + x] = val;
}

LITHO_HD void finalize_pixel(const FinalizeParams& P, int y, int x) {
    float val = 0.f;
    const int yy = y - P.pW, xx = x - P.pW;
    if (yy >= 0 && yy < P.side && xx >= 0 && xx < P.side) {
        if (P.side == P.pn) {  // ATen special case: equal input/output size is a plain copy
            P.out[(size_t)y * P.out_side + x] = fabsf(P.in.at(yy, xx));
            return;
        }
        int r0, r1, c0, c1;
        float lh0, lh1, lw0, lw1;
        bilinear_src(P.scale, yy, P.pn, r0, r1, lh0, lh1);
        bilinear_src(P.scale, xx, P.pn, c0, c1, lw0, lw1);
        const float a = fabsf(P.in.at(r0, c0)), b = fabsf(P.in.at(r0, c1));
        const float c = fabsf(P.in.at(r1, c0)), d = fabsf(P.in.at(r1, c1));
        val = lh0 * (lw0 * a + lw1 * b) + lh1 * (lw0 * c + lw1 * d);
    }
    P.out[(size_t)y * P.out_side + x] = val;
}

}  // namespace litho
