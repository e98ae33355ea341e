"""CPU oracle for the partially coherent Abbe imaging path -- TEST INFRASTRUCTURE ONLY.

This module is a numpy restatement of the reference algorithm
(quarterwave0/LithographySimulator).  It exists so that tests/, bench.py's
``cpu_baseline`` / ``--impl reference`` leg and ``__graft_entry__.smoke()`` can
check the CUDA path.  Nothing under ``lithographysimulator_b200/`` may import it:
the product path has no CPU fallback.

Pinning: the reference ships no tests and no golden vectors, so
the oracle is pinned against outputs of the reference itself, run in the build
container by ``oracle/make_golden.py`` (fixtures committed under tests/golden/);
``tests/test_oracle_golden.py`` re-checks the oracle against those fixtures.

Every function cites the reference file:line it follows.  All builder arithmetic
that the reference does in float16 is restated as "compute in float32, round to
float16 after every op" (``_h``), which is what ATen does on CPU and CUDA.
"""
from __future__ import annotations

import math
import numpy as np

F16 = np.float16
F32 = np.float32


def _h(x):
    """Round to float16 and return as float32 (one fp16 rounding step)."""
    return np.asarray(x, dtype=F32).astype(F16).astype(F32)


# ---------------------------------------------------------------------------------
# Mask.calculateEpsilonN / _nearest2SqInt                      reference mask.py:63-72
# ---------------------------------------------------------------------------------
_POW2 = np.array([2 ** k for k in range(1, 15)], dtype=np.int16)


def nearest_pow2(beta: float) -> int:
    """mask.py:63-65 -- argmin |2^k - beta| over k=1..14 in float32, first wins on ties."""
    d = np.abs(_POW2.astype(F32) - F32(beta))
    return int(_POW2[int(np.argmin(d))])


def calculate_epsilon_n(deltaK: float, pixelSize, wavelength: float):
    """mask.py:67-72 -- beta = wavelength/(deltaK*pixelSize); N = nearest power of two; eps = N/beta."""
    beta = ((deltaK * pixelSize) / wavelength) ** -1
    N = nearest_pow2(beta)
    return N / beta, N


# ---------------------------------------------------------------------------------
# torch.nn.functional.interpolate(mode='bilinear', scale_factor=s, align_corners=False)
# as used at mask.py:77 and imageformation.py:71  
# ---------------------------------------------------------------------------------
def _bilinear_axis(in_size: int, scale_factor: float):
    out_size = int(math.floor(float(in_size) * scale_factor))
    # ATen: coordinate scale is 1/scale_factor cast to float32; the source index is
    # scale*(dst+0.5)-0.5 evaluated in float32 with a fused multiply-add.
    scale = F32(1.0 / scale_factor)
    dst = np.arange(out_size, dtype=np.float64) + 0.5
    src = (np.float64(scale) * dst - 0.5).astype(F32)  # product exact in f64 -> single rounding == FMA
    src = np.maximum(src, F32(0))
    i0 = np.minimum(src.astype(np.int64), in_size - 1)
    i1 = np.minimum(i0 + 1, in_size - 1)
    l1 = (src - i0.astype(F32)).astype(F32)
    l0 = (F32(1) - l1).astype(F32)
    return out_size, i0, i1, l0, l1


def pad2d(img: np.ndarray, lo: int, hi: int) -> np.ndarray:
    """torch.nn.functional.pad(img, (lo, hi, lo, hi)): positive widths add zeros, negative widths crop."""
    n = img.shape[0]
    osz = max(n + lo + hi, 0)
    res = np.zeros((osz, osz), dtype=img.dtype)
    a = max(0, -lo)            # first source index kept
    b = min(n, osz - lo)       # one past the last source index kept
    if b > a:
        res[a + lo:b + lo, a + lo:b + lo] = img[a:b, a:b]
    return res


def bilinear_resize(img: np.ndarray, scale_factor: float, dtype=F32) -> np.ndarray:
    """Square 2-D bilinear resample with torch semantics (same factor on both axes)."""
    img = np.asarray(img, dtype=dtype)
    n = img.shape[0]
    if int(math.floor(float(n) * scale_factor)) == n:
        return img.copy()  # ATen special case: equal input/output size is a plain copy
    _, r0, r1, lh0, lh1 = _bilinear_axis(n, scale_factor)
    _, c0, c1, lw0, lw1 = _bilinear_axis(n, scale_factor)
    lh0 = lh0.astype(dtype)[:, None]; lh1 = lh1.astype(dtype)[:, None]
    lw0 = lw0.astype(dtype)[None, :]; lw1 = lw1.astype(dtype)[None, :]
    a = img[r0][:, c0]; b = img[r0][:, c1]
    c = img[r1][:, c0]; d = img[r1][:, c1]
    return (lh0 * (lw0 * a + lw1 * b) + lh1 * (lw0 * c + lw1 * d)).astype(dtype)


# ---------------------------------------------------------------------------------
# Mask                                                             reference mask.py
# ---------------------------------------------------------------------------------
def demo_geometry() -> np.ndarray:
    """mask.py:22-27 -- the 64x64 four-bar demo mask used when no geometry is given."""
    g = np.zeros((64, 64), dtype=np.int16)
    for c in (16, 25, 34, 43):
        g[9:55, c:c + 4] = 1
    return g


def ff_fraunhofer(geometry: np.ndarray, epsilon: float, N: int, cdtype=np.complex128) -> np.ndarray:
    """mask.py:74-90 -- mask spectrum via the FFT approximation.

    bilinear upsample by epsilon, centre-pad to N, fftshift -> fft2 -> ifftshift, crop pn.
    """
    fdt = F32 if cdtype == np.complex64 else np.float64
    pn = geometry.shape[0]
    scaled = bilinear_resize(geometry.astype(F32), epsilon, dtype=F32).astype(fdt)
    sm = scaled.shape[0]
    pW = ((N - pn) - (sm - pn)) // 2
    corr = sm % 2
    padded = pad2d(scaled, pW, pW + corr)
    spec = np.fft.ifftshift(np.fft.fft2(np.fft.fftshift(padded).astype(cdtype)))
    trim = (N - pn) // 2
    return spec[trim:spec.shape[0] - trim, trim:spec.shape[1] - trim].astype(cdtype)


def _f16_arange(start: float, end: float, step: float) -> np.ndarray:
    """torch.arange(start, end, step, dtype=float16): value_i = fp16(f32(start) + i*f32(step))."""
    n = int(math.ceil((float(end) - float(start)) / float(step)))
    i = np.arange(n, dtype=F32)
    return _h(F32(start) + i * F32(step))


def _direct_phase_table(pn: int, pixelSize) -> np.ndarray:
    """Q[a,c] = fp16(fp16(k[a]) * fp16(x[c]))   imageformation.py:10-24 / mask.py:44-57."""
    deltaK = 4 / pn
    Kbound = pn / 2 * deltaK
    pixelBound = pn / 2 * pixelSize
    k = _f16_arange(-Kbound, Kbound, deltaK)
    x = _f16_arange(-pixelBound, pixelBound, pixelSize)
    return _h(k[:, None] * x[None, :])


def _trapz_weights(n: int, dtype) -> np.ndarray:
    w = np.ones(n, dtype=dtype)
    w[0] = w[-1] = 0.5
    return w


def direct_operator(pn: int, pixelSize, wavelength: float, sign: float, cdtype=np.complex128) -> np.ndarray:
    """A[a,c] = w[c]*exp(sign*i*(2pi/lambda)*Q[a,c]) so that the double trapz integral is A @ G @ A.T.

    The float32 rounding of the argument follows the reference: exponent =
    (q1+q2)*const in complex64; here it is applied per factor, which SURVEY A.2
    measured at 1.3e-7 from the reference.
    """
    Q = _direct_phase_table(pn, pixelSize).astype(np.float64)
    c0 = np.float64(F32(2 * math.pi / wavelength)) if cdtype == np.complex64 else 2 * math.pi / wavelength
    A = np.exp(1j * sign * c0 * Q) * _trapz_weights(pn, np.float64)[None, :]
    return A.astype(cdtype)


def fraunhofer_direct(geometry: np.ndarray, pixelSize, wavelength: float, cdtype=np.complex128) -> np.ndarray:
    """mask.py:41-61 -- direct (trapezoid rule) mask spectrum: A+ @ geometry @ A+^T."""
    pn = geometry.shape[0]
    A = direct_operator(pn, pixelSize, wavelength, +1.0, cdtype)
    return (A @ geometry.astype(cdtype) @ A.T).astype(cdtype)


def fraunhofer(geometry: np.ndarray, pixelSize, wavelength: float, fft: bool, cdtype=np.complex128) -> np.ndarray:
    """mask.py:37-40 dispatch."""
    pn = geometry.shape[0]
    if fft:
        eps, N = calculate_epsilon_n(4 / pn, pixelSize, wavelength)
        return ff_fraunhofer(geometry, eps, N, cdtype)
    return fraunhofer_direct(geometry, pixelSize, wavelength, cdtype)


# ---------------------------------------------------------------------------------
# LightSource                                               reference lightsource.py
# ---------------------------------------------------------------------------------
def _sigma_grid(pn: int, shiftX: float, shiftY: float):
    """lightsource.py:36-46 -- fp16 sigma grid, meshgrid(indexing='xy'), O = sqrt(sX^2+sY^2) in fp16."""
    span = 2
    d = span * 2 / pn
    sx = _f16_arange(-span - shiftX, span - shiftX, d)
    sy = _f16_arange(-span - shiftY, span - shiftY, d)
    sX = np.broadcast_to(sx[None, :], (sy.size, sx.size))
    sY = np.broadcast_to(sy[:, None], (sy.size, sx.size))
    O = _h(np.sqrt(_h(_h(sX * sX) + _h(sY * sY))))
    return sX, sY, O


def light_source_annular(sigmaIn, sigmaOut, pn: int, shiftX=0, shiftY=0) -> np.ndarray:
    """lightsource.py:34-50 -- 1 where sigmaIn <= O <= sigmaOut (thresholds compared in fp16)."""
    _, _, O = _sigma_grid(pn, shiftX, shiftY)
    return ((O >= _h(sigmaIn)) & (O <= _h(sigmaOut))).astype(np.int64)


def light_source_quasar(sigmaIn, sigmaOut, pn: int, count: int, rotation: float, shiftX=0, shiftY=0) -> np.ndarray:
    """lightsource.py:52-73 -- annulus times `count` angular cut-outs, all in fp16."""
    sX, sY, O = _sigma_grid(pn, shiftX, shiftY)
    theta = _h(_h(np.arctan2(sY.astype(np.float64), sX.astype(np.float64))) + F32(rotation))
    two_pi = _h(2 * math.pi)
    theta = _h(theta - two_pi * np.floor(theta / two_pi))  # python-style remainder, fp16 modulus
    theta = np.where(theta == two_pi, F32(0), theta)
    ls = ((O >= _h(sigmaIn)) & (O <= _h(sigmaOut))).astype(np.int64)
    spacing = math.pi / count
    for gap in range(count):
        lo = _h((gap + gap) * spacing)
        hi = _h((gap + gap + 1) * spacing)
        ls = ls * np.where((lo < theta) & (theta < hi), 0, 1)
    return ls


# ---------------------------------------------------------------------------------
# Pupil                                                           reference pupil.py
# ---------------------------------------------------------------------------------
def osa_index_to_mn(ji: int):
    """pupil.py:82-86."""
    n = math.ceil(0.5 * (-3 + math.sqrt(9 + 8 * ji)))
    m = (2 * ji) - (n * (n + 2))
    return m, n


def _pupil_grid(pn: int):
    """pupil.py:50-57 -- fp16 grid, r and theta in fp16."""
    x = _f16_arange(-2, 2, 4 / pn)
    X = np.broadcast_to(x[None, :], (pn, pn))
    Y = np.broadcast_to(x[:, None], (pn, pn))
    r = _h(np.sqrt(_h(_h(X * X) + _h(Y * Y))))
    # transcendental steps are evaluated in float64 and rounded once (float32 libm results differ by an
    # ulp between hosts; the double-rounded value reproduces the reference's CPU tensors exactly)
    theta = _h(np.arctan2(Y.astype(np.float64), X.astype(np.float64)))
    return r, theta


def _pow_f16(r: np.ndarray, e: int) -> np.ndarray:
    """torch.pow(fp16 tensor, python int) -- evaluated in float32, rounded once."""
    if e == 0:
        return np.ones_like(r)
    if e == 1:
        return r.copy()
    if e == 2:
        return _h(r * r)
    if e == 3:
        return _h(r * r * r)
    return _h(np.power(r.astype(np.float64), float(e)))


def generate_z(m: int, n: int, pn: int, coeff_f16: float, grid=None) -> np.ndarray:
    """pupil.py:46-77 -- one OSA Zernike term on the fp16 grid (returns fp16 values as float32)."""
    r, theta = grid if grid is not None else _pupil_grid(pn)
    lLim = int((n - abs(m)) / 2)
    ilLim = int((n + abs(m)) / 2)
    acc = np.zeros((pn, pn), dtype=F32)
    for k in range(lLim + 1):
        static = ((-1) ** k * math.factorial(n - k)) / (
            math.factorial(k) * math.factorial(ilLim - k) * math.factorial(lLim - k))
        acc = acc + _h(F32(static) * _pow_f16(r, n - 2 * k))  # torch.sum over fp16 accumulates in f32
    R = _h(acc)
    Nmn = math.sqrt((2 * n + 1) / (1 + (1 if m == 0 else 0)))
    c = _h(coeff_f16)
    if m >= 0:
        cn = _h(c * F32(Nmn))
        Z = _h(_h(cn * R) * _h(np.cos(_h(F32(m) * theta).astype(np.float64))))
    else:
        cn = _h(c * F32(-Nmn))
        Z = _h(_h(cn * R) * _h(np.sin(_h(F32(m) * theta).astype(np.float64))))
    return np.where(r <= 1, Z, F32(0)).astype(F32)


def wavefront_error(aberrations_f16, pn: int, NA: float, wavelength: float):
    """pupil.py:88-100 -- returns (WE as float32-holding-fp16 values, mutated aberrations).

    Reproduces the in-place defocus rescale aberrations[4] *= NA^2/(4*lambda) (pupil.py:91-92).
    """
    ab = _h(np.array(aberrations_f16, dtype=F32)).copy()
    if len(ab) >= 4:
        ab[4] = _h(_h(ab[4] * F32(NA ** 2)) / F32(4 * wavelength))
    grid = _pupil_grid(pn)
    WE = np.zeros((pn, pn), dtype=F32)
    for i in range(len(ab)):
        m, n = osa_index_to_mn(i)
        WE = _h(WE + generate_z(m, n, pn, ab[i], grid))
    return WE, ab


def pupil_function(aberrations_f16, pn: int, NA: float, wavelength: float, cdtype=np.complex64):
    """pupil.py:32-35,102-111 -- phi = exp(2*pi*i*WE) inside r<=1 (r on the fp16 grid)."""
    WE, ab = wavefront_error(aberrations_f16, pn, NA, wavelength)
    r, _ = _pupil_grid(pn)
    if cdtype == np.complex64:
        arg = (F32(2 * math.pi) * WE).astype(F32).astype(np.float64)
        phi = (np.cos(arg) + 1j * np.sin(arg)).astype(np.complex64)
    else:
        arg = 2 * math.pi * WE.astype(np.float64)
        phi = np.cos(arg) + 1j * np.sin(arg)
    return np.where(r <= 1, phi, 0).astype(cdtype), ab


# ---------------------------------------------------------------------------------
# Image formation                                        reference imageformation.py
# ---------------------------------------------------------------------------------
def source_shifts(lightsource: np.ndarray, pn: int) -> np.ndarray:
    """imageformation.py:59 -- argwhere(lightsource) - pn//2, row-major order, values ignored."""
    return (np.argwhere(lightsource != 0) - (pn // 2)).astype(np.int32)


def calculate_fft_aerial(pf: np.ndarray, maskFT: np.ndarray, pn: int, N: int, cdtype=np.complex128) -> np.ndarray:
    """imageformation.py:32-45 -- product, centre zero-pad to N, fftshift, unnormalised ifft2, ifftshift, crop."""
    prod = pf.astype(cdtype) * maskFT.astype(cdtype)
    pW = (N - pn) // 2
    padded = np.zeros((pn + 2 * pW,) * 2, dtype=cdtype)
    padded[pW:pW + pn, pW:pW + pn] = prod
    field = np.fft.ifft2(np.fft.fftshift(padded), norm="forward")
    field = np.fft.ifftshift(field)
    return field[pW:pW + pn, pW:pW + pn].astype(cdtype)


def calculate_aerial(pupil: np.ndarray, maskFT: np.ndarray, pixelSize, wavelength: float,
                     cdtype=np.complex128, A: np.ndarray | None = None) -> np.ndarray:
    """imageformation.py:3-30 -- direct solver as A- @ (pupil*maskFT) @ A-^T."""
    pn = maskFT.shape[0]
    if A is None:
        A = direct_operator(pn, pixelSize, wavelength, -1.0, cdtype)
    G = pupil.astype(cdtype) * maskFT.astype(cdtype)
    return (A @ G @ A.T).astype(cdtype)


def fft_postprocess(image: np.ndarray, pn: int, epsilon: float, dtype=F32) -> np.ndarray:
    """imageformation.py:69-75 -- abs, bilinear resample by 1/eps, zero border (pW, pW+corr)."""
    image = np.abs(image)
    out = bilinear_resize(image, 1 / epsilon, dtype=dtype)
    pW = (pn - round(pn / epsilon)) // 2
    corr = out.shape[0] % 2
    return pad2d(out, pW, pW + corr)


def abbe_image(maskFT: np.ndarray, pupilF: np.ndarray, lightsource: np.ndarray, pixelSize, deltaK: float,
               wavelength: float, fft: bool, cdtype=np.complex128, shifts: np.ndarray | None = None,
               postprocess: bool = True) -> np.ndarray:
    """imageformation.py:47-77 -- Abbe source-point sum with either solver.

    `cdtype` selects the arithmetic: complex64 follows the reference's precision,
    complex128 is the float64 reference run named by the north star.
    """
    pn = maskFT.shape[0]
    fdt = F32 if cdtype == np.complex64 else np.float64
    if fft:
        eps, N = calculate_epsilon_n(deltaK, pixelSize, wavelength)
    else:
        A = direct_operator(pn, pixelSize, wavelength, -1.0, cdtype)
    if shifts is None:
        shifts = source_shifts(lightsource, pn)
    image = np.zeros((pn, pn), dtype=fdt)
    mft = maskFT.astype(cdtype)
    pup = pupilF.astype(cdtype)
    for d0, d1 in shifts:
        ps = np.roll(pup, (int(d0), int(d1)), axis=(0, 1))  # imageformation.py:63
        if fft:
            e = calculate_fft_aerial(ps, mft, pn, N, cdtype)
        else:
            e = calculate_aerial(ps, mft, pixelSize, wavelength, cdtype, A)
        image += (np.abs(e) ** 2).astype(fdt)
    if fft and postprocess:
        image = fft_postprocess(image, pn, eps, dtype=fdt)
    return image


# ---------------------------------------------------------------------------------
# helpers shared by tests / bench
# ---------------------------------------------------------------------------------
def rel_l2(a: np.ndarray, b: np.ndarray) -> float:
    """SURVEY App. A.6 parity metric: ||a-b||_2 / ||b||_2 (b is the reference)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))
