#!/usr/bin/env python
"""Experiment: does the physical placement of the T ring (a fresh 2.4 GB cudaMalloc per candidate) change the
speed of the accumulation?  Times the same image on several rings kept alive side by side, twice each.
Measurement aid, not part of the product."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from lithographysimulator_b200.imaging import AbbeEngine, source_shifts, epsilon_n  # noqa: E402

dev = torch.device("cuda:0")
cfg, mft, pf, ls = bench.build_inputs_device("cfg3", dev)
eng = AbbeEngine.get(dev)
pn = cfg.pn
eps, N = epsilon_n(4 / pn, cfg.pixel_size, cfg.wavelength)
sh = source_shifts(ls, pn)
plan = eng.plan_for(pn, N, eng.pupil_support(pf), sh)
batch = eng.batch_for(plan, int(sh.shape[0]))
wsb = plan.workspace_bytes(batch)
inten = eng.intensity_plane(plan)
rings, res = [], []
filler = []
for k in range(6):
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    rings.append(ws)
    filler.append(torch.empty((37 + 101 * k) << 20, dtype=torch.uint8, device=dev))   # shift the next ring's placement
for rep in range(2):
    for k, ws in enumerate(rings):
        def run():
            inten.zero_()
            plan.accumulate(mft.data_ptr(), pf.data_ptr(), sh.data_ptr(), None, int(sh.shape[0]), batch, inten.data_ptr(),
                            ws.data_ptr(), wsb, torch.cuda.current_stream(dev).cuda_stream)
        run()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(4):
            run()
        b.record()
        torch.cuda.synchronize()
        res.append({"rep": rep, "ring": k, "ptr": hex(ws.data_ptr()), "ms_per_image": a.elapsed_time(b) / 4})
        print(json.dumps(res[-1]), flush=True)
