#!/usr/bin/env python
"""Library baseline for the hot loop: the reference's op sequence (imageformation.py:32-45, :62-67) written with
stock torch ops -- roll, mul, pad, fftshift, ifft2 (cuFFT on CUDA), ifftshift, crop, abs()**2, += -- on the device
given.  This is what the unmodified reference executes when handed device='cuda' ("the
Blackwell library baseline to beat"); it is a measurement aid, not part of the product and not used by bench.py.

    python scripts/torch_baseline.py [--config cfg3] [--points 64] [--device cuda]

Prints one JSON line: seconds per source point, extrapolated images/s for the whole config, and the relative L2
distance to the product's image for the same source points (when a CUDA device and the library are present).
Without the two implicit .item() syncs per source point that the reference's tensor-valued roll shifts cause."""
import argparse
import json
import math
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lithographysimulator_b200 import workloads as wl  # noqa: E402
from oracle import abbe_oracle as O  # noqa: E402  (input builders only)


def torch_loop(mft, pf, shifts, N):
    pn = mft.shape[0]
    pw = (N - pn) // 2
    image = torch.zeros((pn, pn), dtype=torch.float32, device=mft.device)
    for d0, d1 in shifts:
        g = torch.roll(pf, (int(d0), int(d1)), dims=(0, 1)) * mft                       # imageformation.py:63, :34
        g = torch.nn.functional.pad(g, (pw, pw, pw, pw))                                # :36-37
        e = torch.fft.ifftshift(torch.fft.ifft2(torch.fft.fftshift(g), norm="forward"))  # :39-41
        e = e[pw:pw + pn, pw:pw + pn]                                                   # :43
        image += e.real ** 2 + e.imag ** 2                                              # :67
    return image


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="cfg3")
    ap.add_argument("--points", type=int, default=64)
    ap.add_argument("--device", default="cuda" if torch.cuda.is_available() else "cpu")
    args = ap.parse_args()
    dev = torch.device(args.device)
    cfg = wl.CONFIGS[args.config]
    mft = O.fraunhofer(cfg.geometry(), cfg.pixel_size, cfg.wavelength, True, np.complex64)
    ls = (O.light_source_quasar(cfg.sigma_in, cfg.sigma_out, cfg.pn, 4, -math.pi / 8) if cfg.source == "quasar"
          else O.light_source_annular(cfg.sigma_in, cfg.sigma_out, cfg.pn)) * wl.lattice(cfg.pn, cfg.stride)
    pf, _ = O.pupil_function(cfg.aberrations, cfg.pn, cfg.na, cfg.wavelength)
    shifts_all = O.source_shifts(ls, cfg.pn)
    shifts = shifts_all[:: max(1, len(shifts_all) // args.points)][:args.points]
    _, N = O.calculate_epsilon_n(4 / cfg.pn, cfg.pixel_size, cfg.wavelength)
    mft_d, pf_d = torch.from_numpy(mft).to(dev), torch.from_numpy(pf.astype(np.complex64)).to(dev)

    def sync():
        if dev.type == "cuda":
            torch.cuda.synchronize(dev)

    torch_loop(mft_d, pf_d, shifts[:2], N)   # warm-up (cuFFT plan, allocator)
    sync()
    t0 = time.perf_counter()
    img = torch_loop(mft_d, pf_d, shifts, N)
    sync()
    dt = time.perf_counter() - t0
    out = {"workload": f"{cfg.name}: {cfg.pn}^2, N={N}", "device": str(dev), "points": len(shifts),
           "seconds_per_point": dt / len(shifts), "images_per_s_extrapolated": len(shifts) / dt / len(shifts_all),
           "what": "stock torch ops (roll/mul/pad/fftshift/ifft2/ifftshift/crop/abs^2/+=), one source point at a time"}
    if dev.type == "cuda":
        try:
            from lithographysimulator_b200.imaging import AbbeEngine
            ours = AbbeEngine.get(dev).abbe_fft(mft_d, pf_d, None, cfg.pixel_size, 4 / cfg.pn, cfg.wavelength,
                                                shifts=torch.from_numpy(shifts), postprocess=False)
            out["rel_l2_vs_product"] = float(torch.linalg.vector_norm(ours - img) / torch.linalg.vector_norm(img))
        except Exception as e:  # library not built: the baseline number stands on its own
            out["rel_l2_vs_product"] = f"unavailable: {e}"
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
