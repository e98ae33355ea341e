// Pruned "zoom" DFT bookkeeping shared by all kernels.
//
// A 1-D transform of length L (power of two) is evaluated as
//     X[o'] = sum_u x[u] * exp(s*2*pi*i * o' * c'(u) / L),   o' = o - out.center, o in [0, out.W)
// where the S inputs sit at "true" indices c'(u) = ((in.first + u) mod in.period) - in.center
// (a window of the pn-periodic frequency grid: that is how torch.roll of the pupil,
// imageformation.py:63, appears after folding it into the index arithmetic), and only the W
// centred outputs the reference keeps after its crop (imageformation.py:43) are produced.
//
// Decomposition (R = L/M residues):  o' = R*k' + r  =>
//     X[R*k' + r] = FFT_M( fold_M( x[u] * w_L^(r*c'(u)) ) )[k' mod M]
// fold_M sums inputs whose c' agree modulo M into slot c' mod M.  With M >= S-1 at most two
// inputs share a slot (the pupil rim pixel), so each residue costs one length-M
// FFT instead of a length-L one, and zero inputs are never touched.
//
// Outputs are kept in "residue-major" order: wanted output (r, kk), kk = k' - kmin(r), lives
// at r*Wr + kk.  The column pass and the accumulated intensity use the same ordering on both
// axes, which makes every global access of the two hot kernels unit-stride; the single
// un-permutation happens once per image in the finalize kernel.
#pragma once
#include "hd.h"

namespace litho {

struct AxisIn {
    int first;   // grid index of input 0 before wrapping (any sign)
    int period;  // wrap period (pn); use a large value for "no wrap"
    int center;  // true index = ((first+u) mod period) - center
    int S;       // number of inputs
};

struct AxisOut {
    int W;       // wanted outputs
    int center;  // o' = o - center
};

struct ZoomPlan {
    int L;   // transform length
    int M;   // sub-FFT length
    int R;   // L / M
    int Wr;  // per-residue pitch = max_r cnt(r) = ceil(W / R)
};

// wanted k' range of residue r: o' = R*k' + r in [-center, W-center)
LITHO_HD int zoom_kmin(const AxisOut& o, int R, int r) { return cdiv(-o.center - r, R); }
LITHO_HD int zoom_kend(const AxisOut& o, int R, int r) { return cdiv(o.W - o.center - r, R); }

// natural output index o -> (r, kk)
LITHO_HD void zoom_split(const AxisOut& o, int R, int idx, int& r, int& kk) {
    const int op = idx - o.center;
    r = imod(op, R);
    const int kp = (op - r) / R;  // exact
    kk = kp - zoom_kmin(o, R, r);
}

// Accumulate into `acc` every input u whose true index is congruent to `slot` modulo M,
// multiplied by w_L^(r*c') (conjugated for a forward transform).  ld(u) returns input u.
template <bool FWD, class LoadFn>
LITHO_HD cplx zoom_fold(int slot, int M, int L, int r, const AxisIn& ax, const cplx* twL, LoadFn ld) {
    cplx acc = mk(0.f, 0.f);
    const int f = imod(ax.first, ax.period);
    const int endA = (ax.S < ax.period - f) ? ax.S : (ax.period - f);  // inputs [0,endA) are not wrapped
    {
        const int base = f - ax.center;  // c'(u) = base + u
        for (int u = (slot - base) & (M - 1); u < endA; u += M) {
            cplx x = ld(u);
            if (r != 0) {
                cplx w = ldg_c(twL + ((r * (base + u)) & (L - 1)));
                if (FWD) w = cconj(w);
                x = cmul(x, w);
            }
            acc = cadd(acc, x);
        }
    }
    if (endA < ax.S) {  // wrapped tail: c'(u) = base + u with base shifted by -period
        const int base = f - ax.period - ax.center;
        int u = (slot - base) & (M - 1);
        if (u < endA) u += ((endA - u + M - 1) / M) * M;
        for (; u < ax.S; u += M) {
            cplx x = ld(u);
            if (r != 0) {
                cplx w = ldg_c(twL + ((r * (base + u)) & (L - 1)));
                if (FWD) w = cconj(w);
                x = cmul(x, w);
            }
            acc = cadd(acc, x);
        }
    }
    return acc;
}

// Loads the PPT folded, pre-twiddled inputs of one thread (slots g + TG*e).  When the window does not
// wrap and S <= M+1 (every slot has at most one input, plus the rim input that shares slot of u = 0)
// the loads are issued branch-free and up front, so they are all in flight together; otherwise the
// general two-segment fold is used slot by slot.
template <int PPT, class LoadFn>
LITHO_HD void zoom_load(cplx (&v)[PPT], int g, int TG, int M, int L, int r, const AxisIn& ax, const cplx* twL,
                        LoadFn ld) {
    const int f = imod(ax.first, ax.period);
    const bool simple = (ax.S >= 1) && (f + ax.S <= ax.period) && (ax.S <= M + 1);
    if (simple) {
        const int base = f - ax.center;  // c'(u) = base + u
        const int last = (ax.S < M ? ax.S : M) - 1;
#pragma unroll
        for (int e = 0; e < PPT; ++e) {
            const int u = (g + TG * e - base) & (M - 1);
            v[e] = ld(u <= last ? u : last);
        }
#pragma unroll
        for (int e = 0; e < PPT; ++e) {
            const int u = (g + TG * e - base) & (M - 1);
            cplx x = v[e];
            if (r != 0) x = cmul(x, ldg_c(twL + ((r * (base + u)) & (L - 1))));
            v[e] = mk(u <= last ? x.x : 0.f, u <= last ? x.y : 0.f);
        }
        if (ax.S > M) {  // rim input u = M folds onto the slot of u = 0
            cplx y = ld(M);
            if (r != 0) y = cmul(y, ldg_c(twL + ((r * (base + M)) & (L - 1))));
#pragma unroll
            for (int e = 0; e < PPT; ++e) {
                const bool hit = ((g + TG * e - base) & (M - 1)) == 0;
                v[e] = mk(v[e].x + (hit ? y.x : 0.f), v[e].y + (hit ? y.y : 0.f));
            }
        }
        return;
    }
#pragma unroll
    for (int e = 0; e < PPT; ++e) v[e] = zoom_fold<false>(g + TG * e, M, L, r, ax, twL, ld);
}

}  // namespace litho
