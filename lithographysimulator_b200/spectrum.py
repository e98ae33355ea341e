"""Mask spectrum by the FFT approximation -- reference mask.py:74-90."""
from __future__ import annotations

import torch

from . import _native


def mask_spectrum_fft(geometry: torch.Tensor, epsilon: float, N: int, device: torch.device) -> torch.Tensor:
    """Bilinear upsample by epsilon, centre-pad to N, centred forward DFT, crop to pn x pn (complex64).

    One native call (litho_mask_spectrum): a resample kernel followed by the same pruned zoom-DFT
    row/column kernels the imaging path uses, with a conjugating epilogue for the forward sign.
    """
    lib = _native.device_lib()
    with torch.cuda.device(device):
        geom = geometry.to(device=device, dtype=torch.int16).contiguous()
        pn = int(geom.shape[0])
        nbytes = int(lib.litho_mask_spectrum_workspace_bytes(pn, float(epsilon), int(N)))
        if nbytes == 0:
            raise _native.LithoError(f"mask spectrum: unsupported configuration pn={pn}, eps={epsilon}, N={N}")
        ws = torch.empty(nbytes, dtype=torch.uint8, device=device)
        out = torch.empty((pn, pn), dtype=torch.complex64, device=device)
        lib.check(lib.litho_mask_spectrum(geom.data_ptr(), pn, float(epsilon), int(N), out.data_ptr(), ws.data_ptr(),
                                          nbytes, torch.cuda.current_stream(device).cuda_stream), "litho_mask_spectrum")
        return out
