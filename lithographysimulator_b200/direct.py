"""Direct ("Abbe") solver -- reference imageformation.py:3-30 and mask.py:41-61."""
from __future__ import annotations

from ._native import LithoError


def _missing(*_a, **_k):
    raise LithoError("the direct (fft=False) solver kernels are not built into this library yet; "
                     "there is no CPU fallback")


direct_field = direct_abbe_image = direct_mask_spectrum = _missing
