// In-register DFT butterflies of length 2/4/8/16 (natural order in, natural order out).
//
// X[k] = sum_n a[n] * w^(n k),  w = exp(+2*pi*i/R) for the inverse transform (FWD=false,
// the sign the reference's aerial-image ifft2 uses, imageformation.py:40) and
// exp(-2*pi*i/R) for the forward transform (FWD=true, the mask spectrum fft2, mask.py:84).
// All indices are compile-time constants after unrolling, so the arrays live in registers.
#pragma once
#include "hd.h"

namespace litho {

template <bool FWD>
LITHO_HD cplx mul_i(cplx a) {  // multiply by w4 = +i (inverse) or -i (forward)
    return FWD ? mul_mi(a) : mul_pi(a);
}

// a * (cr + i*s*ci) with s = +1 (inverse) / -1 (forward)
template <bool FWD>
LITHO_HD cplx mul_w(cplx a, float cr, float ci) {
    const float si = FWD ? -ci : ci;
    return mk(a.x * cr - a.y * si, a.x * si + a.y * cr);
}

// Twiddled radix-2 butterfly  (p, m) = e +- w*o,  w = cr + i*s*ci (s = +1 inverse / -1 forward), in
// 6 FMA-pipe instructions instead of 8: with t = si/cr,  w*o = cr*(o.x - t*o.y, o.y + t*o.x), so two
// FMAs form the bracket and four more fold the scale into the add/subtract (Linzer-Feig form; the
// cotangent variant is used when |si| > |cr| so the ratio never exceeds 1).  cr, ci are compile-time
// constants at every call site, so the branch and the division fold away.
template <bool FWD>
LITHO_HD void bfly_w(cplx e, cplx o, float cr, float ci, cplx& p, cplx& m) {
    const float si = FWD ? -ci : ci;
    const float acr = cr < 0.f ? -cr : cr, asi = si < 0.f ? -si : si;
    float dx, dy, sc;
    // explicit fmaf: keeps the compiler from sharing sc*d between the two outputs (mul + 2 adds = 8 ops)
    if (acr >= asi) {
        const float t = si / cr;
        dx = fmaf(-t, o.y, o.x);
        dy = fmaf(t, o.x, o.y);
        sc = cr;
    } else {
        const float t = cr / si;
        dx = fmaf(t, o.x, -o.y);
        dy = fmaf(t, o.y, o.x);
        sc = si;
    }
    p = mk(fmaf(sc, dx, e.x), fmaf(sc, dy, e.y));
    m = mk(fmaf(-sc, dx, e.x), fmaf(-sc, dy, e.y));
}

template <bool FWD>
LITHO_HD void dft2(cplx& a0, cplx& a1) {
    cplx t = a0;
    a0 = cadd(t, a1);
    a1 = csub(t, a1);
}

template <bool FWD>
LITHO_HD void dft4(cplx& a0, cplx& a1, cplx& a2, cplx& a3) {
    cplx b0 = cadd(a0, a2), b1 = csub(a0, a2), b2 = cadd(a1, a3), b3 = mul_i<FWD>(csub(a1, a3));
    a0 = cadd(b0, b2);
    a2 = csub(b0, b2);
    a1 = cadd(b1, b3);
    a3 = csub(b1, b3);
}

template <bool FWD>
LITHO_HD void dft8(cplx (&a)[8]) {
    const float C = 0.70710678118654752440f;
    dft4<FWD>(a[0], a[2], a[4], a[6]);  // E[k] in a[0],a[2],a[4],a[6]
    dft4<FWD>(a[1], a[3], a[5], a[7]);  // O[k] in a[1],a[3],a[5],a[7]
    cplx e0 = a[0], e1 = a[2], e2 = a[4], e3 = a[6];
    cplx o0 = a[1], o1 = a[3], o3 = a[7];
    cplx t2 = mul_i<FWD>(a[5]);
    a[0] = cadd(e0, o0); a[4] = csub(e0, o0);
    bfly_w<FWD>(e1, o1, C, C, a[1], a[5]);
    a[2] = cadd(e2, t2); a[6] = csub(e2, t2);
    bfly_w<FWD>(e3, o3, -C, C, a[3], a[7]);
}

template <bool FWD>
LITHO_HD void dft16(cplx (&a)[16]) {
    const float C1 = 0.92387953251128675613f;  // cos(pi/8)
    const float S1 = 0.38268343236508977173f;  // sin(pi/8)
    const float C2 = 0.70710678118654752440f;  // cos(pi/4)
    cplx e[8], o[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        e[i] = a[2 * i];
        o[i] = a[2 * i + 1];
    }
    dft8<FWD>(e);
    dft8<FWD>(o);
    a[0] = cadd(e[0], o[0]); a[8] = csub(e[0], o[0]);
    bfly_w<FWD>(e[1], o[1], C1, S1, a[1], a[9]);
    bfly_w<FWD>(e[2], o[2], C2, C2, a[2], a[10]);
    bfly_w<FWD>(e[3], o[3], S1, C1, a[3], a[11]);
    {
        const cplx t4 = mul_i<FWD>(o[4]);
        a[4] = cadd(e[4], t4); a[12] = csub(e[4], t4);
    }
    bfly_w<FWD>(e[5], o[5], -S1, C1, a[5], a[13]);
    bfly_w<FWD>(e[6], o[6], -C2, C2, a[6], a[14]);
    bfly_w<FWD>(e[7], o[7], -C1, S1, a[7], a[15]);
}

template <bool FWD>
LITHO_HD void dft32(cplx (&a)[32]) {
    // cos/sin(k*pi/16), k = 0..15
    const float C[16] = {1.0f, 0.98078528040323044913f, 0.92387953251128675613f, 0.83146961230254523708f,
                         0.70710678118654752440f, 0.55557023301960222474f, 0.38268343236508977173f,
                         0.19509032201612826785f, 0.0f, -0.19509032201612826785f, -0.38268343236508977173f,
                         -0.55557023301960222474f, -0.70710678118654752440f, -0.83146961230254523708f,
                         -0.92387953251128675613f, -0.98078528040323044913f};
    const float S[16] = {0.0f, 0.19509032201612826785f, 0.38268343236508977173f, 0.55557023301960222474f,
                         0.70710678118654752440f, 0.83146961230254523708f, 0.92387953251128675613f,
                         0.98078528040323044913f, 1.0f, 0.98078528040323044913f, 0.92387953251128675613f,
                         0.83146961230254523708f, 0.70710678118654752440f, 0.55557023301960222474f,
                         0.38268343236508977173f, 0.19509032201612826785f};
    cplx e[16], o[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        e[i] = a[2 * i];
        o[i] = a[2 * i + 1];
    }
    dft16<FWD>(e);
    dft16<FWD>(o);
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        if (k == 0) {
            a[0] = cadd(e[0], o[0]);
            a[16] = csub(e[0], o[0]);
        } else if (k == 8) {
            const cplx t = mul_i<FWD>(o[8]);
            a[8] = cadd(e[8], t);
            a[24] = csub(e[8], t);
        } else {
            bfly_w<FWD>(e[k], o[k], C[k], S[k], a[k], a[k + 16]);
        }
    }
}

// Generic entry: radix-R DFT over the register subset v[B + t*STRIDE], t = 0..R-1.
template <int R, int STRIDE, int B, bool FWD, int NV>
LITHO_HD void dft_strided(cplx (&v)[NV]) {
    if constexpr (R == 32) {
        cplx a[32];
#pragma unroll
        for (int t = 0; t < 32; ++t) a[t] = v[B + t * STRIDE];
        dft32<FWD>(a);
#pragma unroll
        for (int t = 0; t < 32; ++t) v[B + t * STRIDE] = a[t];
    } else if constexpr (R == 2) {
        dft2<FWD>(v[B], v[B + STRIDE]);
    } else if constexpr (R == 4) {
        dft4<FWD>(v[B], v[B + STRIDE], v[B + 2 * STRIDE], v[B + 3 * STRIDE]);
    } else if constexpr (R == 8) {
        cplx a[8];
#pragma unroll
        for (int t = 0; t < 8; ++t) a[t] = v[B + t * STRIDE];
        dft8<FWD>(a);
#pragma unroll
        for (int t = 0; t < 8; ++t) v[B + t * STRIDE] = a[t];
    } else {
        static_assert(R == 16, "radix must be 2, 4, 8, 16 or 32");
        cplx a[16];
#pragma unroll
        for (int t = 0; t < 16; ++t) a[t] = v[B + t * STRIDE];
        dft16<FWD>(a);
#pragma unroll
        for (int t = 0; t < 16; ++t) v[B + t * STRIDE] = a[t];
    }
}

}  // namespace litho
