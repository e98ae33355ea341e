"""Optical-element objects: ``Mask``, ``LightSource``, ``Pupil`` with the reference's constructor
signatures and attributes (reference mask.py:5, lightsource.py:5, pupil.py:6).

These run once per image, not per source point.  They keep the reference's float16 grid
semantics bit for bit: every elementwise step is evaluated in float32 and
rounded to float16, exactly as ATen does.  Device memory and the run-once elementwise steps go
through torch tensors on the CUDA device; the mask spectrum's transform uses the same native
zoom-DFT kernels as the imaging path.  CUDA devices only -- no CPU fallback.
"""
from __future__ import annotations

import math

import torch

from . import imaging

__all__ = ["Mask", "LightSource", "Pupil", "generateZ", "generateWavefrontError", "generatePhi", "OSAindexToMN", "OSA"]

_SIGMA_SPAN = 2  # grids cover [-2 sigma, +2 sigma): the pupil is the unit disc in the middle half


def _device_or_default(device, what: str) -> torch.device:
    if type(device) is torch.device:
        return imaging._require_cuda(device)
    if not torch.cuda.is_available():
        raise RuntimeError(f"No device defined for {what} and no CUDA device is visible (CUDA only, no CPU fallback)")
    dev = torch.device("cuda", torch.cuda.current_device())
    print(f"No device defined for {what}! Using {torch.cuda.get_device_name(dev)}.")
    return dev


def _half_axis(pn: int, shift, device) -> torch.Tensor:
    """fp16 coordinate axis arange(-2-shift, 2-shift, 4/pn) (lightsource.py:39-40, pupil.py:53)."""
    step = _SIGMA_SPAN * 2 / pn
    return torch.arange(-_SIGMA_SPAN - shift, _SIGMA_SPAN - shift, step, dtype=torch.float16, device=device)


def _radius(ax_x: torch.Tensor, ax_y: torch.Tensor) -> torch.Tensor:
    """fp16 sqrt(x^2 + y^2) on the 'xy' mesh: rows follow y, columns follow x."""
    return torch.sqrt(torch.square(ax_x)[None, :] + torch.square(ax_y)[:, None])


# ------------------------------------------------------------------------------------ Mask
class Mask:
    """Thin binary mask.  Attributes: geometry (int16), pixelNumber, pixelSize, deltaK, device."""

    def __init__(self, geometry: torch.Tensor = None, pixelSize: int = 25, device: torch.device = None):
        self.device = _device_or_default(device, "mask")
        ok = isinstance(geometry, torch.Tensor) and geometry.dim() == 2 and geometry.shape[0] == geometry.shape[1]
        if not ok:
            print("Mask not defined or invalid. Check that it is a torch tensor and is square. Using demo instead.")
            geometry = torch.zeros((64, 64), dtype=torch.int16)
            for c0 in (16, 25, 34, 43):  # the reference's 64x64 four-bar demo pattern (mask.py:22-27)
                geometry[9:55, c0:c0 + 4] = 1
        self.geometry = geometry.to(dtype=torch.int16, device=self.device)
        self.pixelNumber = int(self.geometry.shape[0])
        self.pixelSize = pixelSize
        self.deltaK = 4 / self.pixelNumber
        self._pixelBound = self.pixelNumber / 2 * self.pixelSize
        self._Kbound = self.pixelNumber / 2 * self.deltaK

    def calculateEpsilonN(self, deltaK, pixelSize, wavelength):
        """(epsilon, N) of the FFT approximation -- reference mask.py:63-72."""
        return imaging.epsilon_n(deltaK, pixelSize, wavelength)

    def fraunhofer(self, wavelength, fft: bool) -> torch.Tensor:
        """Mask spectrum: FFT approximation (mask.py:74-90) or direct trapezoid integral (mask.py:41-61)."""
        if fft:
            eps, N = self.calculateEpsilonN(self.deltaK, self.pixelSize, wavelength)
            return self._ffFraunhofer(eps, N)
        from .direct import direct_mask_spectrum
        return direct_mask_spectrum(self.geometry, self.pixelSize, wavelength, self.device)

    def _ffFraunhofer(self, epsilon, N: int) -> torch.Tensor:
        from .spectrum import mask_spectrum_fft
        return mask_spectrum_fft(self.geometry, epsilon, int(N), self.device)


# ----------------------------------------------------------------------------- LightSource
class LightSource:
    """Illumination shapes on the sigma grid (reference lightsource.py:3-73)."""

    def __init__(self, sigmaIn=0, sigmaOut=0.6, pixelNumber: int = 64, NA=0.7, shiftX: int = 0, shiftY: int = 0,
                 device: torch.device = None):
        self.device = _device_or_default(device, "light source")
        self.pixelNumber = pixelNumber
        self.NA = NA
        self.sigmaInner = sigmaIn
        self.sigmaOuter = sigmaOut
        self.shiftX = shiftX
        self.shiftY = shiftY

    def _build(self, count: int, rotation: float) -> torch.Tensor:
        from . import _native
        lib = _native.device_lib()
        pn = int(self.pixelNumber)
        with torch.cuda.device(self.device):
            out = torch.empty((pn, pn), dtype=torch.int64, device=self.device)
            lib.check(lib.litho_source_build(pn, float(self.sigmaInner), float(self.sigmaOuter), float(self.shiftX),
                                             float(self.shiftY), int(count), float(rotation), out.data_ptr(),
                                             torch.cuda.current_stream(self.device).cuda_stream), "litho_source_build")
        return out

    def generateAnnular(self) -> torch.Tensor:
        """1 where sigmaInner <= |sigma| <= sigmaOuter, int64 (lightsource.py:34-50); native kernel."""
        return self._build(0, 0.0)

    def generateQuasar(self, count, rotation) -> torch.Tensor:
        """Annulus with `count` angular gaps removed (lightsource.py:52-73); native kernel for count <= 16."""
        if 1 <= int(count) <= 16:
            return self._build(int(count), float(rotation))
        return quasar_source(self.sigmaInner, self.sigmaOuter, self.pixelNumber, count, rotation, self.shiftX,
                             self.shiftY, self.device)


def annular_source(sigma_in, sigma_out, pn: int, shift_x, shift_y, device) -> torch.Tensor:
    radius = _radius(_half_axis(pn, shift_x, device), _half_axis(pn, shift_y, device))
    return ((radius >= sigma_in) & (radius <= sigma_out)).to(torch.int64)


def quasar_source(sigma_in, sigma_out, pn: int, count, rotation, shift_x, shift_y, device) -> torch.Tensor:
    ax_x, ax_y = _half_axis(pn, shift_x, device), _half_axis(pn, shift_y, device)
    angle = torch.atan2(ax_y[:, None].expand(pn, pn), ax_x[None, :].expand(pn, pn)) + rotation
    angle = torch.remainder(angle, 2 * torch.pi)
    keep = annular_source(sigma_in, sigma_out, pn, shift_x, shift_y, device)
    pitch = torch.pi / count
    for gap in range(count):
        blocked = ((2 * gap) * pitch < angle) & (angle < (2 * gap + 1) * pitch)
        keep = keep * (~blocked).to(torch.int64)
    return keep


# ----------------------------------------------------------------------------------- Pupil
def OSA(m, n):
    return (n * (n + 2) + m) / 2


def OSAindexToMN(ji):
    """OSA/ANSI single index -> (m, n)  (reference pupil.py:82-86)."""
    n = math.ceil((math.sqrt(9 + 8 * ji) - 3) / 2)
    m = 2 * ji - n * (n + 2)
    return m, n


def _polar_grid(pn: int, device):
    ax = _half_axis(pn, 0, device)
    r = _radius(ax, ax)
    theta = torch.arctan2(ax[:, None].expand(pn, pn), ax[None, :].expand(pn, pn))
    return r, theta


def _zernike_term(m: int, n: int, coeff, r: torch.Tensor, theta: torch.Tensor) -> torch.Tensor:
    """coeff * N_mn * R_mn(r) * cos/sin(m theta) in fp16 steps, zero outside the unit disc (pupil.py:46-77)."""
    am = abs(m)
    half_lo, half_hi = (n - am) // 2, (n + am) // 2
    terms = []
    for k in range(half_lo + 1):
        c = ((-1) ** k * math.factorial(n - k)) / (
            math.factorial(k) * math.factorial(half_hi - k) * math.factorial(half_lo - k))
        terms.append(c * r ** (n - 2 * k))
    radial = torch.sum(torch.stack(terms, dim=0), dim=0)  # fp16 in, f32 accumulate, one rounding (as ATen)
    norm = math.sqrt((2 * n + 1) / (2 if m == 0 else 1))
    if m >= 0:
        z = coeff * norm * radial * torch.cos(m * theta)
    else:
        z = coeff * -norm * radial * torch.sin(m * theta)
    return torch.where(r <= 1, z, 0)


def generateZ(m, n, pixelNumber, coeff, device):
    r, theta = _polar_grid(pixelNumber, device)
    return _zernike_term(m, n, coeff, r, theta)


def generateWavefrontError(aberrations, pixelNumber, NA, wavelength, device):
    """Sum of OSA-indexed Zernike terms (pupil.py:88-100).  Like the reference this rescales
    aberrations[4] (defocus, nm) by NA^2/(4 lambda) IN PLACE on the caller's tensor (Q4)."""
    if len(aberrations) >= 4:
        aberrations[4] = aberrations[4] * NA ** 2 / (4 * wavelength)
    return _wavefront_from_scaled(aberrations, pixelNumber, device)


def _wavefront_from_scaled(aberrations, pixelNumber, device):
    r, theta = _polar_grid(pixelNumber, device)
    we = torch.zeros((pixelNumber, pixelNumber), dtype=torch.float16, device=device)
    for j in range(len(aberrations)):
        m, n = OSAindexToMN(j)
        we = we + _zernike_term(m, n, aberrations[j], r, theta)
    return we.type(torch.complex64)


def generatePhi(WE, pixelNumber, device):
    """exp(2 pi i WE) inside the unit disc (pupil.py:102-111)."""
    ax = _half_axis(pixelNumber, 0, device)
    phase = torch.exp(1j * 2 * torch.pi * WE)
    return torch.where(_radius(ax, ax) <= 1, phase, 0)


class Pupil:
    """Projection-lens pupil with Zernike aberrations (reference pupil.py:4-38)."""

    def __init__(self, pixelNumber: int = 64, wavelength=193., NA=0.7, aberrations: torch.Tensor = None,
                 device: torch.device = None):
        self.device = _device_or_default(device, "pupil function")
        if aberrations is None:
            print("No aberrations defined for pupil function! Assuming perfect system.")
            aberrations = torch.tensor([0], dtype=torch.float16, device=self.device)
        self.aberrations = aberrations
        self.pixelNumber = pixelNumber
        self.wavelength = wavelength
        self.NA = NA

    def _build(self, want_pupil: bool, want_we: bool):
        """Native replay of generateWavefrontError + generatePhi (csrc/builders.h).  Like the reference,
        rescales aberrations[4] IN PLACE on the caller's tensor every time it runs (pupil.py:91-92, Q4)."""
        import ctypes as C
        from . import _native
        ab = self.aberrations
        if len(ab) >= 4:
            ab[4] = ab[4] * self.NA ** 2 / (4 * self.wavelength)
        vals = [float(v) for v in ab.detach().to(torch.float16).float().cpu().tolist()]
        if len(vals) > 120:  # radial orders beyond the native table: op-by-op torch evaluation
            we = _wavefront_from_scaled(ab, self.pixelNumber, self.device)
            return (generatePhi(we, self.pixelNumber, self.device) if want_pupil else None), we
        lib = _native.device_lib()
        pn = int(self.pixelNumber)
        with torch.cuda.device(self.device):
            pupil = torch.empty((pn, pn), dtype=torch.complex64, device=self.device) if want_pupil else None
            we = torch.empty((pn, pn), dtype=torch.complex64, device=self.device) if want_we else None
            arr = (C.c_float * len(vals))(*vals)
            lib.check(lib.litho_pupil_build(arr, len(vals), pn, None if pupil is None else pupil.data_ptr(),
                                            None if we is None else we.data_ptr(),
                                            torch.cuda.current_stream(self.device).cuda_stream), "litho_pupil_build")
        return pupil, we

    def generateWavefrontError(self) -> torch.Tensor:
        return self._build(False, True)[1]

    def generatePupilFunction(self) -> torch.Tensor:
        return self._build(True, False)[0]
