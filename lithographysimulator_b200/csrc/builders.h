// Optical-element builders: LightSource (reference lightsource.py:34-73) and Pupil (pupil.py:46-111).
//
// The reference evaluates these on float16 grids: every elementwise torch op computes in float32 and
// rounds its result to float16.  The kernels below replay exactly that op sequence per pixel
// (round_f16 after every step), so the outputs match the reference's CPU tensors bit for bit up to the
// last-ulp behaviour of atan2/cos/sin/pow, which are evaluated in double here and rounded once.
#pragma once
#include "direct_kernels.h"  // round_f16

namespace litho {

LITHO_HD float h16(float v) { return round_f16(v); }

// fp16 coordinate of torch.arange(start, end, step, dtype=float16): fp16(f32(start) + i*f32(step))
LITHO_HD float grid_coord(float start, float step, int i) { return h16(start + (float)i * step); }

LITHO_HD float radius16(float x, float y) { return h16(sqrtf(h16(h16(x * x) + h16(y * y)))); }

LITHO_HD float atan2_16(float y, float x) { return h16((float)atan2((double)y, (double)x)); }

struct SourceParams {
    int pn;
    float x_start, y_start, step;  // float32(-2-shiftX), float32(-2-shiftY), float32(4/pn)
    float sigma_in, sigma_out;     // already rounded to fp16 (comparison operands are cast to the tensor dtype)
    int quasar;                    // 0: annular, 1: quasar
    int count;
    float rotation;                // float32(rotation)
    float two_pi;                  // fp16(2*pi)
    float spacing_lo[16], spacing_hi[16];  // fp16((2g)*pi/count), fp16((2g+1)*pi/count)
    int64_t* out;                  // [pn][pn] int64 0/1
};

LITHO_HD void source_pixel(const SourceParams& P, int i, int j) {
    const float sx = grid_coord(P.x_start, P.step, j);
    const float sy = grid_coord(P.y_start, P.step, i);
    const float o = radius16(sx, sy);
    int64_t v = (o >= P.sigma_in && o <= P.sigma_out) ? 1 : 0;
    if (P.quasar && v) {
        float th = h16(atan2_16(sy, sx) + P.rotation);
        // torch.remainder(theta, fp16(2*pi)): python-style modulo, evaluated in float32, rounded to fp16
        th = h16(th - P.two_pi * floorf(th / P.two_pi));
        if (th == P.two_pi) th = 0.f;
        for (int g = 0; g < P.count; ++g)
            if (P.spacing_lo[g] < th && th < P.spacing_hi[g]) v = 0;
    }
    P.out[(size_t)i * P.pn + j] = v;
}

// One Zernike term: m, n, the fp16 products coeff*(+-N_mn), and its radial polynomial
struct ZernikeTerm {
    int m, n, nk;
    float cn;             // fp16( fp16(coeff) * f32(+N_mn) ) for m >= 0, with -N_mn for m < 0
    float stat[8];        // float32(static coefficient) of r^(n-2k)
    int expo[8];
};

struct PupilParams {
    int pn;
    float start, step;    // float32(-2), float32(4/pn)
    int n_terms;
    const ZernikeTerm* terms;  // device array
    float two_pi_f;       // float32(2*pi)
    cplx* pupil;          // [pn][pn] or null
    cplx* we;             // [pn][pn] wavefront error as complex64 (imag 0) or null
};

LITHO_HD float pow16(float r, int e) {
    if (e == 0) return 1.f;
    if (e == 1) return r;
    if (e == 2) return h16(r * r);
    if (e == 3) return h16(r * r * r);
    return h16((float)pow((double)r, (double)e));
}

LITHO_HD void pupil_pixel(const PupilParams& P, int i, int j) {
    const float x = grid_coord(P.start, P.step, j);
    const float y = grid_coord(P.start, P.step, i);
    const float r = radius16(x, y);
    const float theta = atan2_16(y, x);
    float we = 0.f;
    for (int t = 0; t < P.n_terms; ++t) {
        const ZernikeTerm& z = P.terms[t];
        float acc = 0.f;  // torch.sum over the fp16 stack accumulates in float32, one rounding at the end
        for (int k = 0; k < z.nk; ++k) acc += h16(z.stat[k] * pow16(r, z.expo[k]));
        const float R = h16(acc);
        const float ang = h16((float)z.m * theta);
        const float trig = h16((float)(z.m >= 0 ? cos((double)ang) : sin((double)ang)));
        float zv = h16(h16(z.cn * R) * trig);
        if (!(r <= 1.f)) zv = 0.f;
        we = h16(we + zv);
    }
    const size_t o = (size_t)i * P.pn + j;
    if (P.we) P.we[o] = mk(we, 0.f);
    if (P.pupil) {
        const float arg = P.two_pi_f * we;  // float32 product, as (2*pi*1j) * WE in complex64
        P.pupil[o] = (r <= 1.f) ? mk((float)cos((double)arg), (float)sin((double)arg)) : mk(0.f, 0.f);
    }
}

}  // namespace litho
