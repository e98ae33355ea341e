"""Mask spectrum by the FFT approximation -- reference mask.py:74-90."""
from __future__ import annotations

import torch


def mask_spectrum_fft(geometry: torch.Tensor, epsilon: float, N: int, device: torch.device) -> torch.Tensor:
    """Bilinear upsample by epsilon, centre-pad to N, centred forward DFT, crop to pn x pn (complex64).

    Run-once per mask.  Staging through torch device ops (resample / pad / cuFFT) until the native
    zoom-DFT entry point for real input planes is wired in (kernels exist: ROW_REAL_PLANE).
    """
    pn = int(geometry.shape[0])
    g = geometry.to(device=device, dtype=torch.float32)[None, None]
    scaled = torch.nn.functional.interpolate(g, scale_factor=epsilon, mode="bilinear")[0, 0]
    sm = int(scaled.shape[0])
    lead = ((N - pn) - (sm - pn)) // 2
    trail = lead + sm % 2
    padded = torch.nn.functional.pad(scaled, (lead, trail, lead, trail))
    spec = torch.fft.ifftshift(torch.fft.fft2(torch.fft.fftshift(padded), norm="backward"))
    trim = (N - pn) // 2
    return spec[trim:spec.shape[0] - trim, trim:spec.shape[1] - trim].contiguous()
