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
    cplx t0 = a[1];
    cplx t1 = mul_w<FWD>(a[3], C, C);
    cplx t2 = mul_i<FWD>(a[5]);
    cplx t3 = mul_w<FWD>(a[7], -C, C);
    a[0] = cadd(e0, t0); a[4] = csub(e0, t0);
    a[1] = cadd(e1, t1); a[5] = csub(e1, t1);
    a[2] = cadd(e2, t2); a[6] = csub(e2, t2);
    a[3] = cadd(e3, t3); a[7] = csub(e3, t3);
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
    cplx t[8];
    t[0] = o[0];
    t[1] = mul_w<FWD>(o[1], C1, S1);
    t[2] = mul_w<FWD>(o[2], C2, C2);
    t[3] = mul_w<FWD>(o[3], S1, C1);
    t[4] = mul_i<FWD>(o[4]);
    t[5] = mul_w<FWD>(o[5], -S1, C1);
    t[6] = mul_w<FWD>(o[6], -C2, C2);
    t[7] = mul_w<FWD>(o[7], -C1, S1);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        a[k] = cadd(e[k], t[k]);
        a[k + 8] = csub(e[k], t[k]);
    }
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
        cplx t;
        if (k == 0) t = o[0];
        else if (k == 8) t = mul_i<FWD>(o[8]);
        else t = mul_w<FWD>(o[k], C[k], S[k]);
        a[k] = cadd(e[k], t);
        a[k + 16] = csub(e[k], t);
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
