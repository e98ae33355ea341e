"""Run the UNMODIFIED reference staged under oracle/_ref/ (oracle/build_ref.py) through its own public API.

TEST / BENCH INFRASTRUCTURE ONLY: used by bench.py (--impl reference, cpu_baseline and library_baseline legs)
and by tests; never imported by lithographysimulator_b200.

The only fix-up is the one SURVEY.md App. B-Q1 documents: the reference forgets to import Mask at module scope
(imageformation.py:50 vs :84), so `imageformation.Mask` is injected from its own mask module.  Nothing else is
patched: abbeImage() below is the reference's stock code path (imageformation.py:47-77), on whatever torch
device it is handed ('cpu' = the reference CPU arm, 'cuda' = stock ATen + cuFFT, the library baseline).
"""
from __future__ import annotations

import importlib.util
import math
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
_mods = None


def available() -> bool:
    from oracle import build_ref
    return build_ref.staged()


def load():
    """The four reference modules, imported from oracle/_ref under private names."""
    global _mods
    if _mods is not None:
        return _mods
    if not available():
        raise RuntimeError("oracle/_ref is not staged (run oracle/build_ref.py where /root/reference exists)")
    out = {}
    for name in ("mask", "lightsource", "pupil", "imageformation"):
        spec = importlib.util.spec_from_file_location(f"_litho_ref_{name}", os.path.join(REF_DIR, name + ".py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        out[name] = mod
    out["imageformation"].Mask = out["mask"].Mask      # reference bug Q1
    _mods = out
    return out


def build_inputs(cfg, device):
    """mask -> spectrum -> source -> pupil with the reference's own objects (mask.py, lightsource.py, pupil.py)."""
    import torch
    from lithographysimulator_b200 import workloads as wl
    R = load()
    m = R["mask"].Mask(torch.from_numpy(cfg.geometry()).to(device), cfg.pixel_size, device)
    mft = m.fraunhofer(cfg.wavelength, True)
    src = R["lightsource"].LightSource(cfg.sigma_in, cfg.sigma_out, cfg.pn, cfg.na, 0, 0, device)
    ls = src.generateQuasar(4, -math.pi / 8) if cfg.source == "quasar" else src.generateAnnular()
    ls = ls * torch.from_numpy(wl.lattice(cfg.pn, cfg.stride)).to(device)
    ab = torch.tensor(wl.aberrations_of(cfg), dtype=torch.float16, device=device)
    pf = R["pupil"].Pupil(cfg.pn, cfg.wavelength, cfg.na, ab, device).generatePupilFunction()
    return m, mft, pf, ls


def sub_source(ls, n_points):
    """A light-source tensor holding n_points of ls's source points, evenly spaced in the reference's own
    (row-major argwhere) order: abbeImage() loops over exactly these (imageformation.py:59-62)."""
    import torch
    idx = torch.argwhere(ls)
    n = idx.shape[0]
    step = max(1, n // n_points)
    sel = idx[::step][:n_points]
    out = torch.zeros_like(ls)
    out[sel[:, 0], sel[:, 1]] = 1
    return out, int(sel.shape[0]), int(n)


def abbe_image(m, mft, pf, ls, cfg, device):
    """imageformation.abbeImage(fft=True) exactly as the reference's demo calls it (imageformation.py:119)."""
    R = load()
    return R["imageformation"].abbeImage(m, mft, pf, ls, cfg.pixel_size, m.deltaK, cfg.wavelength, True, device)
