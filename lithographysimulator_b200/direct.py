"""Direct ("Abbe") solver -- reference imageformation.py:3-30 and mask.py:41-61.

E = A * G * A^T with the fp16-quantised phase table, as two native complex matrix products per
source point (csrc/direct_kernels.h).  CUDA only.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _native
from .imaging import AbbeEngine, _as_c64, _check_square, source_shifts


def _operator(lib, pn: int, pixelSize, wavelength: float, sign: int, dev) -> torch.Tensor:
    A = torch.empty((pn, pn), dtype=torch.complex64, device=dev)
    lib.check(lib.litho_direct_operator(pn, float(pixelSize), float(wavelength), sign, A.data_ptr(),
                                        torch.cuda.current_stream(dev).cuda_stream), "litho_direct_operator")
    return A


def _bbox_arg(bbox):
    return (C.c_int * 4)(*bbox)


def direct_abbe_image(maskFT, pupilF, lightsource, pixelSize, wavelength, dev, weights=None, batch: int = 0):
    """abbeImage(fft=False): sum over source points of |A (roll(P) * M) A^T|^2, no post-processing
    (reference imageformation.py:59-65, :77)."""
    eng = AbbeEngine.get(dev)
    lib = eng.lib
    with torch.cuda.device(dev):
        pn = _check_square("maskFT", maskFT)
        _check_square("pupilF", pupilF, pn)
        _check_square("lightsource", lightsource, pn)
        maskFT_d = _as_c64(maskFT, dev)
        pupil_d = _as_c64(pupilF, dev)
        shifts = source_shifts(lightsource.to(dev), pn)
        n_src = int(shifts.shape[0])
        out = torch.zeros((pn, pn), dtype=torch.float32, device=dev)
        if n_src == 0:
            return out
        bbox = eng.pupil_bbox(pupil_d)
        if bbox[1] < bbox[0]:
            return out
        A = _operator(lib, pn, pixelSize, wavelength, -1, dev)
        box = _bbox_arg(bbox)
        if batch <= 0:      # FP32 kernels: 8; tensor-core kernels: enough source points per launch to fill the SMs
            batch = int(lib.litho_direct_default_batch(pn, box))
        batch = max(1, min(batch, n_src))
        nbytes = int(lib.litho_direct_workspace_bytes(pn, box, batch))
        ws = eng.workspace(nbytes)
        w_d = None if weights is None else weights.to(device=dev, dtype=torch.float32).contiguous()
        if w_d is not None and int(w_d.numel()) != n_src:
            raise _native.LithoError(f"{int(w_d.numel())} weights for {n_src} source points")
        lib.check(lib.litho_direct_accumulate(A.data_ptr(), maskFT_d.data_ptr(), pupil_d.data_ptr(), pn, box,
                                              shifts.data_ptr(), None if w_d is None else w_d.data_ptr(), n_src, batch,
                                              out.data_ptr(), ws.data_ptr(), nbytes, eng.stream()),
                  "litho_direct_accumulate")
        return out


def direct_field(pupil, maskFT, fraunhoferConstant, pixelNumber, pixelSize, dev) -> torch.Tensor:
    """calculateAerial: complex field of one (already shifted) pupil.  The wavelength is recovered from
    the reference's `fraunhoferConstant` = -2*pi*i/lambda argument (imageformation.py:52)."""
    eng = AbbeEngine.get(dev)
    lib = eng.lib
    const = complex(fraunhoferConstant)
    if const.imag == 0:
        raise _native.LithoError("calculateAerial: fraunhoferConstant must be +-2*pi*i/wavelength")
    sign = -1 if const.imag < 0 else 1
    wavelength = 2 * torch.pi / abs(const.imag)
    with torch.cuda.device(dev):
        pn = _check_square("maskFT", maskFT)
        _check_square("pupil", pupil, pn)
        maskFT_d = _as_c64(maskFT, dev)
        pupil_d = _as_c64(pupil, dev)
        field = torch.zeros((pn, pn), dtype=torch.complex64, device=dev)
        bbox = eng.pupil_bbox(pupil_d)
        if bbox[1] < bbox[0]:
            return field
        A = _operator(lib, pn, pixelSize, wavelength, sign, dev)
        box = _bbox_arg(bbox)
        nbytes = int(lib.litho_direct_workspace_bytes(pn, box, 1))
        ws = eng.workspace(nbytes)
        lib.check(lib.litho_direct_field(A.data_ptr(), pupil_d.data_ptr(), maskFT_d.data_ptr(), pn, box,
                                         field.data_ptr(), ws.data_ptr(), nbytes, eng.stream()), "litho_direct_field")
        return field


def direct_mask_spectrum(geometry, pixelSize, wavelength, dev) -> torch.Tensor:
    """Mask.fraunhofer(fft=False): A+ * geometry * A+^T (reference mask.py:41-61)."""
    eng = AbbeEngine.get(dev)
    lib = eng.lib
    with torch.cuda.device(dev):
        geom = geometry.to(device=dev, dtype=torch.int16).contiguous()
        pn = int(geom.shape[0])
        A = _operator(lib, pn, pixelSize, wavelength, +1, dev)
        box = _bbox_arg((0, pn - 1, 0, pn - 1))
        nbytes = int(lib.litho_direct_workspace_bytes(pn, box, 1))
        ws = eng.workspace(nbytes)
        out = torch.empty((pn, pn), dtype=torch.complex64, device=dev)
        lib.check(lib.litho_direct_mask_spectrum(A.data_ptr(), geom.data_ptr(), pn, out.data_ptr(), ws.data_ptr(),
                                                 nbytes, eng.stream()), "litho_direct_mask_spectrum")
        return out
