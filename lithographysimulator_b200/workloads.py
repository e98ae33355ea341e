"""Deterministic synthetic workloads for the Abbe imaging path.

Host-side generators only (numpy / CPU torch): masks, lattice-decimated sources and the
five BASELINE.json configurations.  They are shared by tests/, bench.py and the golden
fixture script so that every leg sees bit-identical inputs.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np
import torch

WAVELENGTH = 193.0
PIXEL_SIZE = 25
NA = 0.7
ABERR_FULL = [0, 0, 0.01, 0, 100, 0.01, 0, 0.01, 0.01, 0.01]  # reference imageformation.py:100


def line_space(pn: int) -> np.ndarray:
    """4 px lines on an 8 px pitch (cfg1)."""
    g = np.zeros((pn, pn), dtype=np.int16)
    g[:, (np.arange(pn) % 8) < 4] = 1
    return g


def contacts(pn: int) -> np.ndarray:
    """4x4 px contact holes on a 16 px pitch (cfg2)."""
    a = (np.arange(pn) % 16) < 4
    return (a[:, None] & a[None, :]).astype(np.int16)


def manhattan(pn: int, seed: int = 1234) -> np.ndarray:
    """Random Manhattan rectangles, fill ~0.25 (cfg3-5).  CPU torch generator => platform independent."""
    G = torch.Generator().manual_seed(seed)
    K = pn * pn // 1024
    x0 = torch.randint(0, pn, (K,), generator=G).numpy()
    y0 = torch.randint(0, pn, (K,), generator=G).numpy()
    w = (torch.randint(1, 17, (K,), generator=G) * 2).numpy()
    h = (torch.randint(1, 17, (K,), generator=G) * 2).numpy()
    g = np.zeros((pn, pn), dtype=np.int16)
    for i in range(K):
        g[y0[i]:y0[i] + h[i], x0[i]:x0[i] + w[i]] = 1
    return g


def lattice(pn: int, stride: int) -> np.ndarray:
    """Decimation lattice that always contains the centre pixel."""
    m = np.zeros((pn, pn), dtype=np.int64)
    off = (pn // 2) % stride
    m[off::stride, off::stride] = 1
    return m


@dataclass
class Config:
    name: str
    pn: int
    mask: str                 # line_space | contacts | manhattan
    source: str               # annular | quasar | conventional
    sigma_in: float
    sigma_out: float
    stride: int
    aberrations: list = field(default_factory=lambda: list(ABERR_FULL))
    defocus_sweep: list | None = None   # cfg5: list of defocus values (nm) placed in slot 4
    pixel_size: int = PIXEL_SIZE
    wavelength: float = WAVELENGTH
    na: float = NA

    def geometry(self) -> np.ndarray:
        return {"line_space": line_space, "contacts": contacts, "manhattan": manhattan}[self.mask](self.pn)


def aberrations_of(cfg: "Config", focus_index: int = 0) -> list:
    """Aberration list of one image of the config: cfg5 (a focus sweep) puts defocus_sweep[focus_index] in slot 4
    (the golden image tests/golden/cfg5.npz is the first focus value); the other configs have a single pupil."""
    ab = list(cfg.aberrations)
    if cfg.defocus_sweep:
        ab[4] = cfg.defocus_sweep[focus_index]
    return ab


CONFIGS = {
    "cfg1": Config("cfg1", 256, "line_space", "annular", 0.6, 0.9, 8, aberrations=[0, 0, 0, 0, 50]),
    "cfg2": Config("cfg2", 1024, "contacts", "quasar", 0.4, 0.8, 11),
    "cfg3": Config("cfg3", 2048, "manhattan", "conventional", 0.0, 0.6, 17),
    "cfg4": Config("cfg4", 4096, "manhattan", "annular", 0.6, 0.9, 19),
    "cfg5": Config("cfg5", 8192, "manhattan", "quasar", 0.6, 0.9, 55,
                   defocus_sweep=[float(v) for v in np.linspace(-150, 150, 16)]),
}
