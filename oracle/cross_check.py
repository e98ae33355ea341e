#!/usr/bin/env python
"""Randomised cross-check of the oracle against the UNMODIFIED reference (test infrastructure; needs /root/reference,
so it only runs in the build container).  Each case draws a grid, pixel size, wavelength, NA, source and aberrations,
runs the reference chain (Mask.fraunhofer -> LightSource -> Pupil -> abbeImage, FFT solver) and the oracle chain on
the same parameters, and compares every stage.

    python oracle/cross_check.py [N_CASES] [FIRST_SEED]
"""
import math
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.environ.get("LITHO_REFERENCE", "/root/reference"))
import imageformation as ref_if  # noqa: E402
import lightsource as ref_ls    # noqa: E402
import mask as ref_mask         # noqa: E402
import pupil as ref_pupil       # noqa: E402
ref_if.Mask = ref_mask.Mask

from oracle import abbe_oracle as O  # noqa: E402
from lithographysimulator_b200 import workloads as wl  # noqa: E402

CPU = torch.device("cpu")


def main():
    n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    first = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    worst = {}
    bad = 0
    for seed in range(first, first + n_cases):
        rng = np.random.default_rng(31000 + seed)
        pn = int(rng.choice([32, 64, 128]))            # power-of-two grids: fp16 arange is host-independent there
        ps = int(rng.choice([12, 20, 25, 40, 50]))
        lam = float(rng.choice([193.0, 248.0]))
        na = float(rng.choice([0.5, 0.7, 0.85]))
        eps, N = O.calculate_epsilon_n(4 / pn, ps, lam)
        if N < pn:
            continue
        geom = (rng.random((pn, pn)) < 0.3).astype(np.int16)
        s_in, s_out = sorted(rng.uniform(0.0, 0.95, 2))
        quasar = bool(rng.integers(0, 2))
        stride = int(rng.integers(2, 7))
        n_ab = int(rng.integers(5, 11))
        ab = [float(np.float16(a)) for a in rng.uniform(-0.05, 0.05, n_ab)]
        ab[4] = float(np.float16(rng.uniform(-150, 150)))
        # reference
        m = ref_mask.Mask(torch.from_numpy(geom), ps, CPU)
        mft_r = m.fraunhofer(lam, True)
        L = ref_ls.LightSource(float(s_in), float(s_out), pn, na, 0, 0, CPU)
        ls_r = (L.generateQuasar(4, -math.pi / 8) if quasar else L.generateAnnular()) * torch.from_numpy(wl.lattice(pn, stride))
        pf_r = ref_pupil.Pupil(pn, lam, na, torch.tensor(ab, dtype=torch.float16), CPU).generatePupilFunction()
        if int((ls_r != 0).sum()) == 0:
            continue
        img_r = ref_if.abbeImage(m, mft_r, pf_r, ls_r, ps, m.deltaK, lam, True, CPU).numpy()
        # oracle
        mft_o = O.fraunhofer(geom, ps, lam, True)
        ls_o = (O.light_source_quasar(float(s_in), float(s_out), pn, 4, -math.pi / 8) if quasar
                else O.light_source_annular(float(s_in), float(s_out), pn)) * wl.lattice(pn, stride)
        pf_o, _ = O.pupil_function(ab, pn, na, lam)
        img_o = O.abbe_image(mft_o.astype(np.complex64), pf_o.astype(np.complex64), ls_o, ps, 4 / pn, lam, True, np.complex128)
        errs = {
            "maskFT": float(np.linalg.norm(mft_o - mft_r.numpy()) / np.linalg.norm(mft_r.numpy())),
            "source": float((ls_o != ls_r.numpy()).sum()),
            "pupil": float(np.abs(pf_o - pf_r.numpy()).max()),
            "image": float(O.rel_l2(img_o, img_r)) if img_o.shape == img_r.shape else 1.0,
        }
        for k, v in errs.items():
            worst[k] = max(worst.get(k, 0.0), v)
        if errs["source"] != 0 or errs["image"] > 1e-5 or errs["maskFT"] > 1e-5:
            bad += 1
            print("MISMATCH", seed, dict(pn=pn, ps=ps, lam=lam, na=na, quasar=quasar, sigma=(s_in, s_out)), errs, flush=True)
    print("cases", n_cases, "mismatches", bad, "worst", worst)


if __name__ == "__main__":
    main()
