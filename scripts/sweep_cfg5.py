#!/usr/bin/env python
"""BASELINE cfg5: focus-exposure sweep, 8192^2 mask x 16 defocus values x quadrupole source (980 points),
the 16 pupils sharded over the ranks (lithographysimulator_b200.distributed.focus_sweep_sharded).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/sweep_cfg5.py [--focus 16]

Prints one JSON line (rank 0): seconds per sweep (CUDA events, max over ranks), images/s, fraction of the FP32
roofline with the algorithmic flops of SURVEY.md section 8d.  Inputs are built by the product's own GPU builders."""
import argparse
import json
import math
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lithographysimulator_b200 as L  # noqa: E402
from lithographysimulator_b200 import workloads as wl  # noqa: E402
from lithographysimulator_b200.distributed import focus_sweep_sharded  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--focus", type=int, default=16)
    ap.add_argument("--config", default="cfg5")
    ap.add_argument("--repeats", type=int, default=1)
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cfg = wl.CONFIGS[args.config]
    pn = cfg.pn
    mask = L.Mask(torch.from_numpy(cfg.geometry()), cfg.pixel_size, dev)
    mft = mask.fraunhofer(cfg.wavelength, True)
    src = L.LightSource(cfg.sigma_in, cfg.sigma_out, pn, cfg.na, 0, 0, dev)
    ls = src.generateQuasar(4, -math.pi / 8) if cfg.source == "quasar" else src.generateAnnular()
    ls = ls * torch.from_numpy(wl.lattice(pn, cfg.stride)).to(dev)
    n_src = int((ls != 0).sum())
    sweep = (cfg.defocus_sweep or [0.0])[:args.focus]
    pupils = []
    for d in sweep:
        ab = list(cfg.aberrations)
        ab[4] = d   # (wl.aberrations_of(cfg, i))
        pupils.append(L.Pupil(pn, cfg.wavelength, cfg.na, torch.tensor(ab, dtype=torch.float16, device=dev),
                              dev).generatePupilFunction())
    torch.cuda.synchronize(dev)

    def run():
        return focus_sweep_sharded(mask, mft, pupils, ls, cfg.pixel_size, mask.deltaK, cfg.wavelength, dev)

    imgs = run()  # warm-up (plans, workspaces, NCCL)
    times = []
    for _ in range(args.repeats):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        imgs = run()
        b.record()
        torch.cuda.synchronize(dev)
        t = torch.tensor([a.elapsed_time(b) / 1e3], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        times.append(float(t.item()))
    sec = min(times)
    if rank == 0:
        S, N = pn // 2 + 1, 2 * pn
        w_pt = (S + pn) * 5 * N * math.log2(N) + 6 * S * S + 4 * pn * pn
        flops = w_pt * n_src * len(sweep)
        finite = all(bool(torch.isfinite(i).all()) for i in imgs)
        parity = None
        gpath = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", f"{cfg.name}.npz")
        if os.path.exists(gpath) and cfg.defocus_sweep:      # the golden is the sweep's first focus value
            z = np.load(gpath)
            st = int(z["sample_stride"])
            ref = torch.from_numpy(z["image_sample"]).to(dev).double()
            parity = float((imgs[0][::st, ::st].double() - ref).norm() / ref.norm())
        print(json.dumps({"workload": f"{cfg.name}: {pn}^2 mask x {len(sweep)} defocus values x {n_src} source points",
                          "n_gpus": world, "seconds_per_sweep": sec, "images_per_s": len(sweep) / sec,
                          "algorithmic_tflops": flops / sec / 1e12, "images": len(imgs),
                          "image_side": int(imgs[0].shape[0]), "finite": finite,
                          "rel_l2_focus0_vs_reference_golden": parity,
                          "sharding": "pupils (focus values) over ranks, no reduce; a rank's focus values share one row "
                                      "pass per batch of source points (focus batching); images all-gathered"}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
