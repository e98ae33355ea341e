// Shared-memory Stockham FFT of compile-time length M, 16 points per thread.
//
// One FFT is carried out by a group of TG = M/PPT threads.  Thread g (0 <= g < TG) holds
// element  g + TG*e  in register e, both on entry (natural-order input) and on exit
// (natural-order output).  The passes use radices {16,16,...,rest}; between passes the data
// is exchanged through shared memory with a 1-in-16 padding that removes the bank
// conflicts of the stride-16 scatter of the first pass.
//
// `es` is the shared-memory element stride: 1 when the group owns a private region (row
// mode) and CB when CB adjacent columns are interleaved (column mode, column fastest).
#pragma once
#include "butterfly.h"

namespace litho {

LITHO_HD int pad16(int i) { return i + (i >> 4); }

template <int M>
struct FftShape {
    static_assert(M >= 16 && (M & (M - 1)) == 0 && M <= 16384, "M must be a power of two in [16,16384]");
    static constexpr int PPT = 16;
    static constexpr int TG = M / PPT;
    static constexpr int SMEM_ELEMS = M + M / 16;  // padded element count per FFT
    // radices: as many 16s as fit, then the remainder
    static constexpr int log2M() {
        int l = 0;
        for (int m = M; m > 1; m >>= 1) ++l;
        return l;
    }
    static constexpr int NP = (log2M() + 3) / 4;
    static constexpr int radix(int pass) {
        int rem = log2M() - 4 * pass;
        return rem >= 4 ? 16 : (1 << rem);
    }
};

// One Stockham pass.  PASS is the pass index, NS the product of the previous radices.
template <int M, int PASS, int NS, bool FWD, class Ctx>
LITHO_HD void fft_pass(cplx (&v)[16], cplx* sm, int es, int g, const cplx* tw, int twscale, const Ctx& ctx) {
    using Sh = FftShape<M>;
    constexpr int R = Sh::radix(PASS);
    constexpr int NB = 16 / R;  // butterflies per thread in this pass
    constexpr int TG = Sh::TG;
    constexpr bool LAST = (PASS == Sh::NP - 1);

    if constexpr (PASS > 0) {
        // gather this pass's inputs: element g + TG*e -> register e
        ctx.sync();
#pragma unroll
        for (int e = 0; e < 16; ++e) v[e] = sm[pad16(g + TG * e) * es];
        // twiddle: butterfly b has index j = g + b*TG, k = j mod NS, angle 2*pi*t*k/(NS*R)
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            const int j = g + b * TG;
            const int k = j & (NS - 1);
            const int step = k * (M / (NS * R)) * twscale;  // index of w_M^(k*M/(NS*R)) in the w_L table
#pragma unroll
            for (int t = 1; t < R; ++t) {
                cplx w = ldg_c(tw + t * step);
                if (FWD) w = cconj(w);
                v[b + t * NB] = cmul(v[b + t * NB], w);
            }
        }
    }

    // NB independent radix-R butterflies over registers {b + t*NB}
    if constexpr (NB == 1) {
        dft_strided<R, 1, 0, FWD>(v);
    } else {
        // unrolled over b with compile-time register indices
        if constexpr (NB >= 2) { dft_strided<R, NB, 0, FWD>(v); dft_strided<R, NB, 1, FWD>(v); }
        if constexpr (NB >= 4) { dft_strided<R, NB, 2, FWD>(v); dft_strided<R, NB, 3, FWD>(v); }
        if constexpr (NB >= 8) {
            dft_strided<R, NB, 4, FWD>(v); dft_strided<R, NB, 5, FWD>(v);
            dft_strided<R, NB, 6, FWD>(v); dft_strided<R, NB, 7, FWD>(v);
        }
    }

    if constexpr (!LAST) {
        // scatter: output t of butterfly j goes to (j-k)*R + k + t*NS
        ctx.sync();
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            const int j = g + b * TG;
            const int k = j & (NS - 1);
            const int base = (j - k) * R + k;
#pragma unroll
            for (int t = 0; t < R; ++t) sm[pad16(base + t * NS) * es] = v[b + t * NB];
        }
        fft_pass<M, PASS + 1, NS * R, FWD>(v, sm, es, g, tw, twscale, ctx);
    }
    // LAST: j < M/R = NS so k = j and the output index is g + TG*(b + t*NB): register e holds
    // element g + TG*e again.
}

// Full transform.  `tw` is the table w_L[i] = exp(+2*pi*i*i/L) (L = M*twscale entries).
// Every thread of the CTA must call this the same number of times (it contains CTA-wide syncs).
template <int M, bool FWD, class Ctx>
LITHO_HD void fft_run(cplx (&v)[16], cplx* sm, int es, int g, const cplx* tw, int twscale, const Ctx& ctx) {
    fft_pass<M, 0, 1, FWD>(v, sm, es, g, tw, twscale, ctx);
}

}  // namespace litho
