// Shared-memory Stockham FFT of compile-time length M with PPT (16 or 32) points per thread.
//
// One FFT is carried out by a group of TG = M/PPT threads.  Thread g (0 <= g < TG) holds
// element  g + TG*e  in register e, both on entry (natural-order input) and on exit
// (natural-order output).  The passes use radices {PPT, PPT, ..., rest}; between passes the data
// is exchanged through shared memory with a 1-in-PPT padding that removes the bank conflicts of
// the stride-PPT scatter of the first pass.
//
// `es` is the shared-memory element stride: 1 when the group owns a private region (row mode)
// and CB when CB adjacent columns are interleaved (column mode, column fastest).
//
// Two policies are passed in:
//   Tw   : tw.get<PASS>(t, k) returns w^(t*k) with w = exp(+2*pi*i/(NS*R)) of that pass
//          (global w_L table for the generic kernels, per-pass shared-memory tables for the fast ones);
//   Sync : sync() separates the shared-memory phases of one FFT group (CTA barrier, or a warp
//          barrier when the whole group lives in one warp).
#pragma once
#include "butterfly.h"

namespace litho {

template <int PPT>
LITHO_HD int padp(int i) {
    return i + (i / PPT);
}

template <int M, int PPT = 16>
struct FftShape {
    static_assert(M >= PPT && (M & (M - 1)) == 0 && M <= 16384, "M must be a power of two in [PPT,16384]");
    static_assert(PPT == 16 || PPT == 32, "PPT must be 16 or 32");
    static constexpr int LOGP = (PPT == 16) ? 4 : 5;
    static constexpr int TG = M / PPT;
    static constexpr int SMEM_ELEMS = M + M / PPT;  // padded element count per FFT
    static constexpr int log2M() {
        int l = 0;
        for (int m = M; m > 1; m >>= 1) ++l;
        return l;
    }
    static constexpr int NP = (log2M() + LOGP - 1) / LOGP;
    static constexpr int radix(int pass) {
        int rem = log2M() - LOGP * pass;
        return rem >= LOGP ? PPT : (1 << rem);
    }
    // product of the radices before `pass`
    static constexpr int ns(int pass) {
        int n = 1;
        for (int p = 0; p < pass; ++p) n *= radix(p);
        return n;
    }
};

template <int R, int NB, int B, bool FWD, int PPT>
LITHO_HD void dft_all(cplx (&v)[PPT]) {
    dft_strided<R, NB, B, FWD>(v);
    if constexpr (B + 1 < NB) dft_all<R, NB, B + 1, FWD>(v);
}

// One Stockham pass.  PASS is the pass index, NS the product of the previous radices.
// Hook: hook.after_last_gather() runs once per FFT right after the last read of the exchange buffer
// (before the last butterflies), i.e. at the point from which the buffer is free again -- the fast
// column kernel uses it to start the asynchronous copy of its next input tile into that buffer.
// hook.after_first_sync() runs once per multi-pass FFT right after the CTA/group barrier that precedes the first
// scatter, i.e. when every thread of the group has consumed its inputs -- the TMA-staged column kernel uses it
// to start the copy of the next input tile.
struct NoHook {
    LITHO_HD void after_last_gather() const {}
    LITHO_HD void after_first_sync() const {}
};

template <int M, int PPT, int PASS, bool FWD, class Tw, class Sync, class Hook>
LITHO_HD void fft_pass(cplx (&v)[PPT], cplx* sm, int es, int g, const Tw& tw, const Sync& sync, const Hook& hook) {
    using Sh = FftShape<M, PPT>;
    constexpr int R = Sh::radix(PASS);
    constexpr int NS = Sh::ns(PASS);
    constexpr int NB = PPT / R;  // butterflies per thread in this pass
    constexpr int TG = Sh::TG;
    constexpr bool LAST = (PASS == Sh::NP - 1);

    if constexpr (PASS == 0 && LAST) hook.after_last_gather();  // single pass: the buffer is never used
    if constexpr (PASS > 0) {
        // gather this pass's inputs: element g + TG*e -> register e
        sync.sync();
        if constexpr (TG % PPT == 0) {
            // padp(g + TG*e) = (g + g/PPT) + e*(TG + TG/PPT) exactly when PPT divides TG: one per-thread base and
            // compile-time offsets (left to itself the compiler rebuilds the padding with a mask per element)
            const cplx* gp = sm + (g + g / PPT) * es;
#pragma unroll
            for (int e = 0; e < PPT; ++e) v[e] = gp[e * (TG + TG / PPT) * es];
        } else {
#pragma unroll
            for (int e = 0; e < PPT; ++e) v[e] = sm[padp<PPT>(g + TG * e) * es];
        }
        if constexpr (LAST) hook.after_last_gather();
        // twiddle: butterfly b has index j = g + b*TG, k = j mod NS, angle 2*pi*t*k/(NS*R)
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            const int k = (g + b * TG) & (NS - 1);
#pragma unroll
            for (int t = 1; t < R; ++t) {
                cplx w = tw.template get<PASS>(t, k);
                if (FWD) w = cconj(w);
                v[b + t * NB] = cmul(v[b + t * NB], w);
            }
        }
    }

    // NB independent radix-R butterflies over registers {b + t*NB}
    dft_all<R, NB, 0, FWD>(v);

    if constexpr (!LAST) {
        // scatter: output t of butterfly j goes to (j-k)*R + k + t*NS
        sync.sync();
        if constexpr (PASS == 0) hook.after_first_sync();
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            const int j = g + b * TG;
            const int k = j & (NS - 1);
            const int base = (j - k) * R + k;
#pragma unroll
            for (int t = 0; t < R; ++t) sm[padp<PPT>(base + t * NS) * es] = v[b + t * NB];
        }
        fft_pass<M, PPT, PASS + 1, FWD>(v, sm, es, g, tw, sync, hook);
    }
    // LAST: j < M/R = NS so k = j and the output index is g + TG*(b + t*NB): register e holds
    // element g + TG*e again.
}

// Full transform.  Every thread of the sync scope must call this the same number of times.
template <int M, int PPT, bool FWD, class Tw, class Sync, class Hook = NoHook>
LITHO_HD void fft_run(cplx (&v)[PPT], cplx* sm, int es, int g, const Tw& tw, const Sync& sync,
                      const Hook& hook = Hook()) {
    LITHO_ASSUME(((unsigned)g < (unsigned)FftShape<M, PPT>::TG));
    fft_pass<M, PPT, 0, FWD>(v, sm, es, g, tw, sync, hook);
}

// Twiddles from the global table w_L[i] = exp(+2*pi*i*i/L), L = M*twscale (generic kernels).
template <int M, int PPT>
struct GlobalTw {
    const cplx* tw;
    int twscale;
    template <int PASS>
    LITHO_HD cplx get(int t, int k) const {
        using Sh = FftShape<M, PPT>;
        constexpr int step = M / (Sh::ns(PASS) * Sh::radix(PASS));
        return ldg_c(tw + t * k * step * twscale);
    }
};

template <class Ctx>
struct CtaSync {
    const Ctx& ctx;
    LITHO_HD void sync() const { ctx.sync(); }
};

}  // namespace litho
