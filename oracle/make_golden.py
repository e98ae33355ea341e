#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ by running the UNMODIFIED reference.

TEST INFRASTRUCTURE ONLY.  Runs in the build container, where /root/reference is mounted
(it does not exist on the GPU box).  The reference has no tests or golden vectors of its
own, so these outputs of the reference itself are the parity pins:

  tests/golden/kat_small.npz   full inputs + outputs of small cases (64..128 px)
  tests/golden/cfg1.npz        BASELINE cfg1 (256 px) inputs + full image
  tests/golden/cfg2.npz        BASELINE cfg2 (1024 px): strided image sample + moments
  tests/golden/cfg3.npz        BASELINE cfg3 (2048 px): strided image sample + moments
  tests/golden/cfg4_subset.npz BASELINE cfg4 grid (4096 px, N = 8192): 8 of its 4104 source points
  tests/golden/cfg5_subset.npz BASELINE cfg5 grid (8192 px, N = 16384, the reference's own 4099-px-wide pupil at
                               defocus -150 nm): 3 of its 980 source points
                               (the full configs would take the reference hours: SURVEY section 8c)

  tests/golden/cfg4.npz, cfg5.npz  the FULL cfg4 (4104 points) and cfg5 (980 points, first focus value) images:
                               about an hour of CPU each, not part of the default list

Usage:  python oracle/make_golden.py [kat cfg1 cfg2 cfg3 cfg4_subset cfg5_subset cfg4 cfg5]
"""
from __future__ import annotations

import math
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("LITHO_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)

import imageformation as ref_if  # noqa: E402  (the reference)
import lightsource as ref_ls    # noqa: E402
import mask as ref_mask         # noqa: E402
import pupil as ref_pupil       # noqa: E402

ref_if.Mask = ref_mask.Mask  # reference bug Q1: Mask is only imported under __main__ (imageformation.py:50 vs :84)

from lithographysimulator_b200 import workloads as wl  # noqa: E402

CPU = torch.device("cpu")
OUT = os.path.join(ROOT, "tests", "golden")
SAMPLE = 16  # stride of the image sample kept for the large configs


def ref_case(geometry, pixel_size, source, pupil_ab, fft, wavelength=193.0, na=0.7, pupil_tensor=None,
             ls_tensor=None):
    """Run the reference chain mask -> spectrum -> source -> pupil -> abbeImage on CPU."""
    g = None if geometry is None else torch.from_numpy(np.asarray(geometry))
    m = ref_mask.Mask(g, pixel_size, CPU)
    pn = m.pixelNumber
    mft = m.fraunhofer(wavelength, fft)
    if ls_tensor is None:
        kind, s_in, s_out, stride, sx, sy = source
        L = ref_ls.LightSource(s_in, s_out, pn, na, sx, sy, CPU)
        ls = L.generateAnnular() if kind in ("annular", "conventional") else L.generateQuasar(4, -math.pi / 8)
        if stride > 1:
            ls = ls * torch.from_numpy(wl.lattice(pn, stride))
    else:
        ls = torch.from_numpy(ls_tensor)
    if pupil_tensor is None:
        ab = torch.tensor(pupil_ab, dtype=torch.float16)
        pf = ref_pupil.Pupil(pn, wavelength, na, ab, CPU).generatePupilFunction()
    else:
        pf = torch.from_numpy(pupil_tensor)
    eps, N = m.calculateEpsilonN(m.deltaK, pixel_size, wavelength)
    t = time.time()
    img = ref_if.abbeImage(m, mft, pf, ls, pixel_size, m.deltaK, wavelength, fft, CPU)
    dt = time.time() - t
    return dict(geometry=m.geometry.numpy(), maskFT=mft.numpy(), pupil=pf.numpy(), lightsource=ls.numpy(),
                image=img.numpy(), eps=np.float64(eps), N=np.int64(N), pixel_size=np.int64(pixel_size),
                fft=np.bool_(fft), seconds=np.float64(dt))


def pack(prefix, d, out):
    for k, v in d.items():
        out[f"{prefix}/{k}"] = v


def make_kat():
    out = {}
    ab10 = list(wl.ABERR_FULL)
    # reference demo (imageformation.py:99-119) with both source shapes
    pack("demo64_quasar", ref_case(None, 25, ("quasar", 0.4, 0.8, 1, 0, 0), ab10, True), out)
    pack("demo64_annular", ref_case(None, 25, ("annular", 0.4, 0.8, 1, 0, 0), ab10, True), out)
    # N == pn (pixelSize 50) and N == 4 pn (pixelSize 12)
    pack("ps50_64", ref_case(None, 50, ("annular", 0.3, 0.7, 2, 0, 0), ab10, True), out)
    pack("ps12_64", ref_case(None, 12, ("quasar", 0.4, 0.8, 3, 0, 0), ab10, True), out)
    # even, non power-of-two grid
    pack("np2_96", ref_case(wl.contacts(96), 25, ("annular", 0.5, 0.9, 5, 0, 0), [0, 0, 0, 0, 50], True), out)
    # source reaching beyond the pupil edge: roll() wraps around the grid
    pack("wrap_128", ref_case(wl.manhattan(128, 7), 25, ("annular", 0.3, 1.3, 9, 0, 0), ab10, True), out)
    pack("shifted_128", ref_case(wl.manhattan(128, 7), 25, ("quasar", 0.4, 0.8, 7, 0.5, -0.25), ab10, True), out)
    # arbitrary dense complex pupil (no zero support at all) and weighted (non 0/1) source values
    rng = np.random.default_rng(5)
    dense = (rng.standard_normal((64, 64)) + 1j * rng.standard_normal((64, 64))).astype(np.complex64)
    lsw = (wl.lattice(64, 5) * rng.integers(1, 9, (64, 64))).astype(np.int64)
    pack("dense_64", ref_case(None, 25, None, None, True, pupil_tensor=dense, ls_tensor=lsw), out)
    # direct ("Abbe") solver, a handful of source points (0.6 s/pt at 64, 7 s/pt and 8 GB at 128)
    ls6 = np.zeros((64, 64), dtype=np.int64)
    for r, c in ((32, 32), (20, 32), (32, 45), (40, 24), (25, 25), (44, 41)):
        ls6[r, c] = 1
    pack("direct_64", ref_case(None, 25, None, ab10, False, ls_tensor=ls6), out)
    # (power-of-two grids only: for other sizes ATen's CPU float16 arange is vector-width dependent,
    #  so the direct solver's fp16 coordinate grids would differ between hosts and from CUDA ATen)
    ls2 = np.zeros((128, 128), dtype=np.int64)
    for r, c in ((64, 64), (40, 70)):
        ls2[r, c] = 1
    pack("direct_128", ref_case(wl.line_space(128), 25, None, [0, 0, 0, 0, 50], False, ls_tensor=ls2), out)
    # single-field entry points
    m = ref_mask.Mask(None, 25, CPU)
    mft = m.fraunhofer(193.0, True)
    pf = ref_pupil.Pupil(64, 193.0, 0.7, torch.tensor(ab10, dtype=torch.float16), CPU).generatePupilFunction()
    pfs = torch.roll(pf, shifts=(5, -9), dims=(0, 1))
    out["field_fft_64/pf"] = pfs.numpy()
    out["field_fft_64/maskFT"] = mft.numpy()
    out["field_fft_64/field"] = ref_if.calculateFFTAerial(pfs, mft, 64, 128).numpy()
    mftd = m.fraunhofer(193.0, False)
    out["field_direct_64/pf"] = pfs.numpy()
    out["field_direct_64/maskFT"] = mftd.numpy()
    out["field_direct_64/field"] = ref_if.calculateAerial(pfs, mftd, (-2 * 1j * torch.pi) / 193.0, 64, 25, CPU).numpy()
    np.savez_compressed(os.path.join(OUT, "kat_small.npz"), **out)
    print("kat_small written", len(out), "arrays")


def make_cfg(name: str, full_image: bool, sample: int = SAMPLE):
    """A whole BASELINE config through the unmodified reference (cfg5: its first focus value, -150 nm)."""
    cfg = wl.CONFIGS[name]
    t0 = time.time()
    kind = cfg.source
    ab = list(cfg.aberrations)
    if cfg.defocus_sweep:
        ab[4] = cfg.defocus_sweep[0]
    d = ref_case(cfg.geometry(), cfg.pixel_size, (kind, cfg.sigma_in, cfg.sigma_out, cfg.stride, 0, 0),
                 ab, True, cfg.wavelength, cfg.na)
    img = d["image"]
    out = dict(eps=d["eps"], N=d["N"], seconds=d["seconds"], n_src=np.int64((d["lightsource"] != 0).sum()),
               aberrations=np.array(ab),
               shape=np.array(img.shape), img_sum=np.float64(img.sum(dtype=np.float64)),
               img_sumsq=np.float64((img.astype(np.float64) ** 2).sum()), img_max=np.float64(img.max()),
               sample_stride=np.int64(sample), image_sample=img[::sample, ::sample].copy(),
               # spot checks that pin the (builder) inputs too
               maskFT_abs_sum=np.float64(np.abs(d["maskFT"]).sum(dtype=np.float64)),
               maskFT_sample=d["maskFT"][::sample, ::sample].copy(),
               pupil_nnz=np.int64((d["pupil"] != 0).sum()), pupil_sample=d["pupil"][::sample, ::sample].copy(),
               ls_rows=np.argwhere(d["lightsource"] != 0).astype(np.int32))
    if full_image:
        out.update(image=img, maskFT=d["maskFT"], pupil=d["pupil"])
    np.savez_compressed(os.path.join(OUT, f"{name}.npz"), **out)
    print(name, "written; reference abbeImage", float(d["seconds"]), "s; total", time.time() - t0, "s")


def make_cfg_subset(name: str, n_points: int, sample: int):
    """A few source points of a large config through the unmodified reference (its loop is strictly per source
    point, imageformation.py:62-67, so a subset of the source is a valid `lightsource` argument)."""
    cfg = wl.CONFIGS[name]
    pn = cfg.pn
    t0 = time.time()
    L = ref_ls.LightSource(cfg.sigma_in, cfg.sigma_out, pn, cfg.na, 0, 0, CPU)
    ls = L.generateAnnular() if cfg.source in ("annular", "conventional") else L.generateQuasar(4, -math.pi / 8)
    ls = (ls * torch.from_numpy(wl.lattice(pn, cfg.stride))).numpy()
    rows = np.argwhere(ls != 0)
    pick = rows[np.linspace(0, len(rows) - 1, n_points).round().astype(int)]   # spread over the source, extremes included
    sub = np.zeros_like(ls)
    sub[pick[:, 0], pick[:, 1]] = 1
    ab = list(cfg.aberrations)
    if cfg.defocus_sweep:
        ab[4] = cfg.defocus_sweep[0]
    d = ref_case(cfg.geometry(), cfg.pixel_size, None, ab, True, cfg.wavelength, cfg.na, ls_tensor=sub)
    img = d["image"]
    nz = np.argwhere(d["pupil"] != 0)
    out = dict(eps=d["eps"], N=d["N"], seconds=d["seconds"], n_src_full=np.int64(len(rows)), aberrations=np.array(ab),
               shape=np.array(img.shape), img_sum=np.float64(img.sum(dtype=np.float64)),
               img_sumsq=np.float64((img.astype(np.float64) ** 2).sum()), img_max=np.float64(img.max()),
               sample_stride=np.int64(sample), image_sample=img[::sample, ::sample].copy(),
               maskFT_abs_sum=np.float64(np.abs(d["maskFT"]).sum(dtype=np.float64)),
               maskFT_sample=d["maskFT"][::sample, ::sample].copy(),
               pupil_nnz=np.int64(len(nz)), pupil_bbox=np.array([nz[:, 0].min(), nz[:, 0].max(), nz[:, 1].min(), nz[:, 1].max()]),
               pupil_sample=d["pupil"][::sample, ::sample].copy(), ls_rows=pick.astype(np.int32))
    np.savez_compressed(os.path.join(OUT, f"{name}_subset.npz"), **out)
    print(name, "subset written;", n_points, "points; reference abbeImage", float(d["seconds"]), "s; total",
          time.time() - t0, "s; image", img.shape, "pupil bbox", out["pupil_bbox"])


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    torch.manual_seed(0)
    what = sys.argv[1:] or ["kat", "cfg1", "cfg2", "cfg3"]
    for w in what:
        if w == "kat":
            make_kat()
        elif w == "cfg4_subset":
            make_cfg_subset("cfg4", 8, 16)
        elif w == "cfg5_subset":
            make_cfg_subset("cfg5", 3, 32)
        else:
            make_cfg(w, full_image=(w == "cfg1"), sample=32 if w == "cfg5" else SAMPLE)
