#!/usr/bin/env python
"""Headline benchmark: aerial images/s at 2048^2 x ~1k source points (BASELINE.json cfg3).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config cfg3]

One "step" = one full aerial image: abbeImage(fft=True) over all source points of the config
(mask spectrum and pupil are inputs, as in the reference's call).  Prints ONE JSON line.

  value        images/s with inputs resident in HBM, CUDA events around the K timed images, L2 flushed between
               images, max over ranks.  N > 1: the source points are sharded across ranks; the rank that
               post-processes image i (i mod N) sums the partial intensity planes of all ranks with loads over
               NVLink from their CUDA-IPC-mapped buffers (litho_peer_sum; --reduce nccl: one ncclReduce instead),
               so the job is ONE image computed N-way ("scaling": "strong").
  e2e          same metric through the public API with HOST (pinned) tensors: H2D of mask spectrum, pupil and
               source, source-point extraction, compute, D2H of the image, all inside the timed region; the
               copies of image i+1 overlap the kernels of image i (AbbeEngine.prepare / ShardedPipeline.submit).
  roofline     dominant kernel (column pass) timed alone with CUDA events on its stream; algorithmic
               flops per SURVEY.md section 8d; FP32 peak measured in this run by an FMA probe.
  parity       rel-L2 of the timed loop's last image against the unmodified reference's image of the same config
               (tests/golden/<cfg>.npz).
  cpu_baseline the UNMODIFIED reference (oracle/_ref, staged by oracle/build_ref.py) abbeImage(fft=True) on CPU
               tensors, all host threads, on two sub-sources of the config (slope = cost per source point, the
               loop is strictly per source point, reference imageformation.py:62-67); the numpy port of the
               oracle only where oracle/_ref is not staged.
  library_baseline  the same unmodified reference handed device='cuda': stock ATen + cuFFT on the same B200.
  --impl reference  the reference arm: that CPU run alone, same config string, JSON line with "impl": "reference".
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "aerial images/s at 2048^2 x 1k source pts"
UNIT = "images/s"


# ----------------------------------------------------------------------------- workload
def build_inputs_device(cfg_name: str, dev):
    """Synthetic inputs of a BASELINE config built by the product's own builders on the GPU (native
    kernels behind Mask.fraunhofer / LightSource / Pupil) -- the oracle is not involved in this arm."""
    import torch
    import lithographysimulator_b200 as L
    from lithographysimulator_b200 import workloads as wl
    cfg = wl.CONFIGS[cfg_name]
    mask = L.Mask(torch.from_numpy(cfg.geometry()), cfg.pixel_size, dev)
    mft = mask.fraunhofer(cfg.wavelength, True)
    src = L.LightSource(cfg.sigma_in, cfg.sigma_out, cfg.pn, cfg.na, 0, 0, dev)
    ls = src.generateQuasar(4, -math.pi / 8) if cfg.source == "quasar" else src.generateAnnular()
    ls = ls * torch.from_numpy(wl.lattice(cfg.pn, cfg.stride)).to(dev)
    ab = torch.tensor(wl.aberrations_of(cfg), dtype=torch.float16, device=dev)
    pf = L.Pupil(cfg.pn, cfg.wavelength, cfg.na, ab, dev).generatePupilFunction()
    torch.cuda.synchronize(dev)
    return cfg, mft, pf, ls


def build_inputs_host(cfg_name: str):
    """Synthetic inputs of a BASELINE config, built on the host with the oracle builders (numpy); used by
    the reference arm and the cpu_baseline leg only."""
    from oracle import abbe_oracle as O
    from lithographysimulator_b200 import workloads as wl
    cfg = wl.CONFIGS[cfg_name]
    geom = cfg.geometry()
    mft = O.fraunhofer(geom, cfg.pixel_size, cfg.wavelength, True, np.complex128).astype(np.complex64)
    if cfg.source == "quasar":
        ls = O.light_source_quasar(cfg.sigma_in, cfg.sigma_out, cfg.pn, 4, -math.pi / 8)
    else:
        ls = O.light_source_annular(cfg.sigma_in, cfg.sigma_out, cfg.pn)
    ls = (ls * wl.lattice(cfg.pn, cfg.stride)).astype(np.int64)
    pf, _ = O.pupil_function(wl.aberrations_of(cfg), cfg.pn, cfg.na, cfg.wavelength)
    return cfg, mft, pf.astype(np.complex64), ls


def workload_string(cfg, n_src: int, N: int) -> str:
    """One string for both arms (the driver compares config.workload of the two lines)."""
    return (f"{cfg.name}: {cfg.pn}^2 {cfg.mask} mask, {cfg.source} source {n_src} pts, N={N}, "
            "Zernike-aberrated pupil, FFT-approximation solver")


def algorithmic_flops(pn: int, N: int, n_src: int):
    """W_pt = (S + pn)*5*N*log2(N) + 6*S^2 + 4*pn^2, S = pn/2+1."""
    S = pn // 2 + 1
    lg = math.log2(N)
    rows = S * 5 * N * lg + 6 * S * S
    cols = pn * 5 * N * lg + 4 * pn * pn
    return dict(rows=rows * n_src, cols=cols * n_src, total=(rows + cols) * n_src, per_point=rows + cols)


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    """SM clock and throttle reasons sampled while the benchmark runs: NVML polled from a thread every 10 ms
    (a timed region can be shorter than nvidia-smi's 100 ms loop), nvidia-smi as the fallback.  mark() brackets
    the timed region; stop() reports the samples taken inside it (all samples if none fell inside)."""
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []
        self.samples = []      # (t, sm_mhz, max_mhz, set of reason names)
        self.marks = []
        self.mode = None
        self._stop = threading.Event()

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.gpu)
            mx = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            get_reasons = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
                pynvml.nvmlDeviceGetCurrentClocksThrottleReasons
            bits = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}

            def poll():
                while not self._stop.is_set():
                    try:
                        sm = float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                        r = int(get_reasons(h))
                        self.samples.append((time.perf_counter(), sm, mx, {n for n, b in bits.items() if r & b}))
                    except Exception:
                        pass
                    self._stop.wait(0.01)

            self.thread = threading.Thread(target=poll, daemon=True)
            self.thread.start()
            self.mode = "nvml"
            return
        except Exception:
            self.mode = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
            self.mode = "nvidia-smi"
        except Exception:
            self.proc = None

    def mark(self):
        self.marks.append(time.perf_counter())

    def _read(self):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.proc.stdout:
            f = [x.strip() for x in line.strip().split(",")]
            if len(f) < 9:
                continue
            try:
                sm, mx = float(f[1]), float(f[2])
            except ValueError:
                continue
            self.samples.append((time.perf_counter(), sm, mx,
                                 {nm for nm, val in zip(names, f[5:9]) if val.lower().startswith("active")}))

    def stop(self):
        if self.mode is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no NVML / nvidia-smi"]}
        self._stop.set()
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()
        self.thread.join(timeout=2)
        inside = self.samples
        if len(self.marks) >= 2:
            t0, t1 = self.marks[0], self.marks[1]
            sel = [x for x in self.samples if t0 <= x[0] <= t1]
            if sel:
                inside = sel
        sm = [x[1] for x in inside]
        mx = [x[2] for x in inside]
        reasons = set().union(*[x[3] for x in inside]) if inside else set()
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "samples_total": len(self.samples), "source": self.mode,
                "window": "timed region" if inside is not self.samples else "whole run",
                "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------- reference arm
def oracle_image_sample(mft, pf, shifts, N, threads):
    """|E_s|^2 summed over `shifts` with the oracle's FFT solver, source points spread over host threads."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import abbe_oracle as O
    pn = mft.shape[0]

    def work(chunk):
        acc = np.zeros((pn, pn), dtype=np.float32)
        for d0, d1 in chunk:
            e = O.calculate_fft_aerial(np.roll(pf, (int(d0), int(d1)), axis=(0, 1)), mft, pn, N, np.complex64)
            acc += (e.real ** 2 + e.imag ** 2)
        return acc

    chunks = [shifts[i::threads] for i in range(threads) if len(shifts[i::threads])]
    with ThreadPoolExecutor(max_workers=threads) as ex:
        parts = list(ex.map(work, chunks))
    return sum(parts)


def time_cpu_baseline(cfg, mft, pf, ls, n_sample, threads):
    from oracle import abbe_oracle as O
    shifts = O.source_shifts(ls, cfg.pn)
    n_src = len(shifts)
    _, N = O.calculate_epsilon_n(4 / cfg.pn, cfg.pixel_size, cfg.wavelength)
    sample = shifts[:: max(1, n_src // n_sample)][:n_sample]
    oracle_image_sample(mft, pf, sample[:threads], N, threads)  # warm-up (thread pool, FFT plans)
    t0 = time.perf_counter()
    oracle_image_sample(mft, pf, sample, N, threads)
    dt = time.perf_counter() - t0
    per_image = dt / len(sample) * n_src
    return 1.0 / per_image, len(sample), n_src, dt


def time_reference(cfg, device, n_lo, n_hi, reps, budget_s, threads=None):
    """The UNMODIFIED reference (oracle/_ref, staged by oracle/build_ref.py) through its own public API:
    mask.Mask.fraunhofer, lightsource.LightSource, pupil.Pupil, imageformation.abbeImage(fft=True) on `device`.
    One "step" times abbeImage() on two sub-sources of the config's source (n_lo and n_hi of its points, evenly
    spaced in the reference's own loop order): the slope is the cost per source point, the intercept the
    once-per-image work (argwhere, post-processing), and one image = intercept + slope * n_src -- the loop is
    strictly per source point (imageformation.py:62-67).  Returns (images/s per step, n_src, N, seconds spent)."""
    import torch
    from oracle import ref_runner as RR
    if threads:
        torch.set_num_threads(threads)
    dev = torch.device(device)
    m, mft, pf, ls = RR.build_inputs(cfg, dev)
    _, N = m.calculateEpsilonN(m.deltaK, cfg.pixel_size, cfg.wavelength)
    ls_lo, n_lo, n_src = RR.sub_source(ls, n_lo)
    ls_hi, n_hi, _ = RR.sub_source(ls, n_hi)

    def run(src):
        if dev.type == "cuda":
            torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        img = RR.abbe_image(m, mft, pf, src, cfg, dev)
        if dev.type == "cuda":
            torch.cuda.synchronize(dev)
        return time.perf_counter() - t0, img

    run(ls_lo)                       # discarded warm-up (thread pool, FFT plans; 5-40x slower)
    vals, spent, img = [], 0.0, None
    for _ in range(max(1, reps)):
        t_lo, _ = run(ls_lo)
        t_hi, img = run(ls_hi)
        slope = max((t_hi - t_lo) / max(1, n_hi - n_lo), 1e-9)
        icpt = max(t_lo - slope * n_lo, 0.0)
        vals.append(1.0 / (icpt + slope * n_src))
        spent += t_lo + t_hi
        if spent > budget_s:
            break
    return vals, n_src, int(N), spent, (n_lo, n_hi), (ls_hi, img)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from lithographysimulator_b200 import workloads as wl
    from oracle import ref_runner as RR
    cfg = wl.CONFIGS[args.config]
    threads = os.cpu_count() or 1
    if RR.available():
        n_hi = max(8, args.ref_sample)
        vals, n_src, N, spent, (n_lo, n_hi), _ = time_reference(cfg, "cpu", max(2, n_hi // 4), n_hi, args.steps, 170.0, threads)
        kind = "reference"
        sample = (f"unmodified reference (oracle/_ref) abbeImage(fft=True, device=cpu), torch {threads} threads: per step "
                  f"two sub-sources of {n_lo} and {n_hi} of the {n_src} source points, image time = intercept + slope*n_src; "
                  f"{len(vals)} steps, {spent:.1f} s of CPU work")
    else:
        cfg, mft, pf, ls = build_inputs_host(args.config)
        from oracle import abbe_oracle as O
        _, N = O.calculate_epsilon_n(4 / cfg.pn, cfg.pixel_size, cfg.wavelength)
        n_sample = max(threads, min(args.ref_sample, 4 * threads))
        vals = []
        for _ in range(max(1, min(args.steps, 8))):
            v, ns, n_src, dt = time_cpu_baseline(cfg, mft, pf, ls, n_sample, threads)
            vals.append(v)
        kind = "port"
        sample = (f"oracle/_ref not staged on this box: oracle numpy port (pocketfft complex64), one source point per host "
                  f"thread, {ns} of {n_src} source points per step, extrapolated linearly in n_src")
    v = float(np.median(vals))
    line = {"metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": len(vals), "warmup": 1,
            "ms_per_step": 1000.0 / v, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "complex64", "data": "synthetic", "impl": "reference",
            "config": {"workload": workload_string(cfg, n_src, N)},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- our arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    import lithographysimulator_b200 as L
    from lithographysimulator_b200 import _native
    from lithographysimulator_b200.imaging import AbbeEngine, source_shifts, epsilon_n

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _native.device_lib()

    cfg, mft_d, pf_d, ls_d = build_inputs_device(args.config, dev)
    pn = cfg.pn
    eps, N = epsilon_n(4 / pn, cfg.pixel_size, cfg.wavelength)
    mft_h, pf_h, ls_h = mft_d.cpu().numpy(), pf_d.cpu().numpy(), ls_d.cpu().numpy()
    eng = AbbeEngine.get(dev)
    shifts_all = source_shifts(ls_d, pn)
    n_src = int(shifts_all.shape[0])
    shifts_mine = shifts_all[rank::world].contiguous()  # interleaved shard: equal work per rank
    support = eng.pupil_support(pf_d)
    # one plan for all ranks, chosen from ALL source points (the partial planes are summed)
    plan = eng.plan(pn, N, support, generic=True) if args.generic else eng.plan_for(pn, N, support, shifts_all)
    flush = torch.empty(160 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    # Pipelined, sharded imager (lithographysimulator_b200.distributed.ShardedPipeline): image i is accumulated by
    # all ranks together; rank i mod N sums the partial planes -- over peer memory (NVLink loads, litho_peer_sum)
    # unless --reduce nccl -- and alone post-processes it on a second stream while every rank already accumulates
    # image i+1.  All buffers of a step are allocated up front.
    from lithographysimulator_b200.distributed import ShardedPipeline
    pipe = ShardedPipeline(eng, plan, eps, reduce=args.reduce)
    fin_stream = pipe.fin_stream

    def one_image(prep=None, out_host=None):
        """One aerial image.  prep = None: inputs resident in HBM (`value`); otherwise a PreparedImage staged from
        pinned host tensors by eng.prepare() on the copy stream (`e2e`), and the root copies the image to out_host."""
        if args.no_pipeline:
            inten = pipe.planes[0]
            inten.zero_()
            if prep is None:
                eng.accumulate(plan, mft_d, pf_d, shifts_mine, inten, None, args.batch)
            else:
                torch.cuda.current_stream(dev).wait_event(prep.ready)
                eng.accumulate(prep.plan, prep.maskFT, prep.pupil, prep.shifts, inten, None, args.batch)
                eng.consumed(prep)
            if world > 1:
                dist.all_reduce(inten)
            pipe.last_image = eng.finalize(plan, inten, eps)
            if out_host is not None:
                out_host.copy_(pipe.last_image, non_blocking=True)
            return
        if prep is None:
            pipe.submit(mft_d, pf_d, shifts_mine, batch=args.batch, inputs_ready=not args.no_chain)
        else:
            prep.shifts.record_stream(torch.cuda.current_stream(dev))
            pipe.submit(prep.maskFT, prep.pupil, prep.shifts, batch=args.batch, wait_event=prep.ready,
                        on_accumulated=lambda: eng.consumed(prep), out_host=out_host)

    def join():
        pipe.join()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    warm = max(args.warmup, 3)   # timing rules: at least 3 warm-up steps
    for _ in range(max(warm, world)):   # ... and every rank is the root of at least one warm-up image
        one_image()
    join()
    barrier()
    pipe.check()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev_a, ev_b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    sampler.mark()
    t_wall0 = time.perf_counter()
    ev_a.record()
    for _ in range(args.steps):
        flush.fill_(1)          # evict L2 between images (inside the timed region: 256 MB write, ~0.05 ms)
        one_image()
    join()
    ev_b.record()
    barrier()
    sampler.mark()
    t_wall = time.perf_counter() - t_wall0
    ms = ev_a.elapsed_time(ev_b)
    tms = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms = float(tms.item())
    ms_per_step = ms / args.steps
    value = 1000.0 / ms_per_step

    # ---- parity of the images the timed loop produced against the unmodified reference's image of this config
    # (tests/golden/<cfg>.npz, oracle/make_golden.py: every 16th pixel of the reference's CPU result) ----
    parity = None
    gpath = os.path.join(ROOT, "tests", "golden", f"{cfg.name}.npz")
    if os.path.exists(gpath):
        pv = torch.tensor([-1.0], dtype=torch.float64, device=dev)
        z = np.load(gpath)
        st = int(z["sample_stride"])
        ref_s = torch.from_numpy(z["image_sample"]).to(dev)
        if pipe.last_image is not None and tuple(pipe.last_image[::st, ::st].shape) == tuple(ref_s.shape):
            pv[0] = (pipe.last_image[::st, ::st].double() - ref_s.double()).norm() / ref_s.double().norm()
        if world > 1:
            dist.all_reduce(pv, op=dist.ReduceOp.MAX)   # every rank that post-processed a timed image reports
        if float(pv.item()) >= 0:
            parity = {"rel_l2_vs_reference_golden": float(pv.item()), "tolerance": 1e-5,
                      "golden": f"tests/golden/{cfg.name}.npz (unmodified reference on CPU, every {st}th pixel)",
                      "image": "last image(s) of the timed loop (chained, pipelined path)"}
            assert parity["rel_l2_vs_reference_golden"] < 1e-5, parity

    # ---- optional per-image event trace of an extra, untimed loop (all ranks; --trace) ----
    trace = None
    if args.trace:
        barrier()
        pipe.trace = []
        for _ in range(args.steps):
            flush.fill_(1)
            one_image()
        join()
        barrier()
        mine_tr = pipe.trace_ms()
        pipe.trace = None
        if world > 1:
            allt = [None] * world
            dist.all_gather_object(allt, mine_tr)
        else:
            allt = [mine_tr]
        if rank == 0:
            acc = [[t["accumulated"] - t["begin"] for t in r] for r in allt]
            roots = [t for r in allt for t in r if "finalized" in t]
            trace = {"accumulate_ms_per_rank_mean": [float(np.mean(a)) for a in acc],
                     "accumulate_ms_max": float(np.max(acc)), "accumulate_ms_min": float(np.min(acc)),
                     "image_period_ms_per_rank": [float((r[-1]["begin"] - r[0]["begin"]) / max(1, len(r) - 1)) for r in allt],
                     "root_wait_for_peers_plus_sum_ms": [round(t["summed"] - t["fin_begin"], 4) for t in roots],
                     "root_finalize_ms": [round(t["finalized"] - t["summed"], 4) for t in roots],
                     "root_lag_fin_begin_after_accumulated_ms": [round(t["fin_begin"] - t["accumulated"], 4) for t in roots]}

    # ---- phase breakdown of one image (events on the current stream, after the timed region) ----
    def timed(fn, reps=3):
        fn()                            # untimed first call: workspaces of this code path get allocated here
        torch.cuda.synchronize(dev)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize(dev)
        return a.elapsed_time(b) / reps

    inten_b = eng.intensity_plane(plan)
    breakdown = {
        "zero_plane": timed(lambda: inten_b.zero_()),
        "accumulate": timed(lambda: eng.accumulate(plan, mft_d, pf_d, shifts_mine, inten_b, None, args.batch)),
        "finalize": timed(lambda: eng.finalize(plan, inten_b, eps)),
    }

    # ---- per-kernel timing (live, CUDA events on the launching stream, after the timed region) ----
    inten = eng.intensity_plane(plan)
    n_mine = int(shifts_mine.shape[0])
    batch = eng.batch_for(plan, n_mine, args.batch)
    wsb = plan.workspace_bytes(batch)
    ws = eng.workspace(wsb)
    kern = {}
    for name, phases in (("rows", 1), ("cols", 2)):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(dev)
        a.record()
        plan.accumulate(mft_d.data_ptr(), pf_d.data_ptr(), shifts_mine.data_ptr(), None, n_mine, batch,
                        inten.data_ptr(), ws.data_ptr(), wsb, eng.stream(), phases=phases)
        b.record()
        torch.cuda.synchronize(dev)
        launches = (n_mine + batch - 1) // batch
        kern[name] = {"ms_total": a.elapsed_time(b), "launches": launches}

    # ---- FP32 peak measured in this run (FMA probe) ----
    probe_out = torch.zeros(4, dtype=torch.float32, device=dev)
    import ctypes as C
    fl = C.c_double()
    nsm = torch.cuda.get_device_properties(dev).multi_processor_count
    best = 0.0
    for _ in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        lib.check(lib.litho_fp32_probe(probe_out.data_ptr(), nsm * 8, 20000, C.byref(fl), eng.stream()), "probe")
        b.record()
        torch.cuda.synchronize(dev)
        best = max(best, fl.value / (a.elapsed_time(b) * 1e-3) / 1e12)
    clocks = sampler.stop() if rank == 0 else None

    # ---- end to end through the public API with host (pinned) tensors ----
    # Every step: H2D of that image's mask spectrum, pupil and source (full [pn,pn] tensors, as the reference's
    # abbeImage takes them) from pinned memory, source-point extraction, compute, D2H of the finished image.
    # eng.prepare() stages image i+1 on the copy stream while image i is computed (two staging sets), and the
    # D2H runs on the post-processing stream, so the copies overlap the kernels but are all inside the timed region.
    mft_p = torch.from_numpy(mft_h).pin_memory()
    pf_p = torch.from_numpy(pf_h).pin_memory()
    ls_p = torch.from_numpy(ls_h).pin_memory()
    side = plan.output_side(eps)
    out_p = [torch.empty((side, side), dtype=torch.float32).pin_memory() for _ in range(2)]
    shard = (rank, world) if world > 1 else None
    # N > 1: every rank uploads 1/N of each input tensor and the slices are all-gathered over NVLink on a
    # communicator of their own (so the gathers do not queue behind the intensity reduce)
    # ... or (default) pulled from the peers' staging buffers by the copy engines (distributed.PeerStaging): no NCCL
    # kernels next to the persistent compute kernels
    upload_pg, upload_peers = None, None
    if world > 1 and args.upload == "peer":
        from lithographysimulator_b200.distributed import PeerStaging, PeerUnavailable

        def exchange(obj):
            out = [None] * world
            dist.all_gather_object(out, obj)
            return out
        try:
            upload_peers = PeerStaging(eng.lib, AbbeEngine.peer_staging_bytes(pn, ls_p.dtype), rank, world, exchange)
        except PeerUnavailable:          # raised on every rank: all fall back to the NCCL all-gather together
            args.upload = "nccl"
        torch.cuda.synchronize(dev)
    if world > 1 and args.upload == "nccl":
        upload_pg = dist.new_group(backend="nccl")

    def prepare(i):
        return eng.prepare(mft_p, pf_p, ls_p, cfg.pixel_size, 4 / pn, cfg.wavelength, slot=i % 2, shard=shard,
                           plan=plan, upload_group=upload_pg, upload_peers=upload_peers)

    def e2e_loop(n):
        prep = prepare(0)
        for i in range(n):
            one_image(prep, out_p[i % 2])
            prep = prepare(i + 1) if i + 1 < n else None    # staged while image i is being computed
        join()
        torch.cuda.synchronize(dev)

    e2e_loop(2 * max(1, min(world, 8)))   # warm-up: staging sets, the upload communicator, every root (even count:
    barrier()                             # the image counter keeps its slot parity)
    e2e_steps = max(2, min(args.steps, 8))
    t0 = time.perf_counter()
    e2e_loop(e2e_steps)
    barrier()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_s = float(te.item())
    h2d = mft_p.numel() * 8 + pf_p.numel() * 8 + ls_p.numel() * 8   # whole job: the slices of all ranks add up to this
    if world > 1 and upload_pg is None and upload_peers is None:
        h2d *= world                                                 # every rank pulls the full tensors
    d2h = out_p[0].numel() * 4
    # the image that came back over PCIe must be the image the resident-input path produced
    e2e_check = None
    if world == 1 and pipe.last_image is not None:
        ref_img = eng.abbe_fft(mft_d, pf_d, ls_d, cfg.pixel_size, 4 / pn, cfg.wavelength, plan=plan).double()
        got = out_p[(e2e_steps - 1) % 2].to(dev).double()     # float64 norms: the raw intensities are ~1e17
        e2e_check = float((got - ref_img).norm() / ref_img.norm())
    pipe.check()

    if rank == 0:
        fl_alg = algorithmic_flops(pn, N, n_mine)
        cols_ms = kern["cols"]["ms_total"] / kern["cols"]["launches"]
        cols_flops = fl_alg["cols"] / kern["cols"]["launches"]
        achieved = cols_flops / (cols_ms * 1e-3) / 1e12
        peak_nominal = nsm * 128 * 2 * (clocks["sm_max_mhz"] or 1965.0) * 1e6 / 1e12
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        step_ach = algorithmic_flops(pn, N, n_src)["total"] / (ms_per_step * 1e-3) / 1e12
        # DRAM bytes per launch of the dominant kernel from the committed ncu capture (scaled by batch)
        traffic, traffic_src, dram_per_image = None, None, None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            if tj.get("config") == cfg.name and plan.path == 2:
                traffic = (tj["dram_bytes_read_per_launch"] + tj["dram_bytes_write_per_launch"]) * batch / tj["batch"]
                traffic_src = tj["source"]
                rk = tj.get("rows_kernel")
                if rk:   # both hot kernels, scaled from the captured batch to the whole image's source points
                    per_pt = (tj["dram_bytes_read_per_launch"] + tj["dram_bytes_write_per_launch"] +
                              rk["dram_bytes_read_per_launch"] + rk["dram_bytes_write_per_launch"]) / tj["batch"]
                    dram_per_image = per_pt * n_src
        except Exception:
            pass
        roofline = {
            "bound": "fp32", "kernel": (("abbe_fast_cols_tma_kernel" if plan.column_tile() > 0 else "abbe_fast_cols_kernel")
                                        if plan.path == 2 else "abbe_cols_kernel") +
                      " (column pass + |E|^2 accumulate)",
            "achieved": achieved, "peak": best, "unit": "TFLOP/s", "frac": achieved / best if best else None,
            "frac_of_nominal_peak": achieved / peak_nominal,
            "peak_source": "FP32 FMA probe measured in this run (MEASURED_PEAKS.json has no FP32 entry); "
                           f"nominal {peak_nominal:.1f} TFLOP/s = SMs*128*2*max clock",
            "flops_per_launch": cols_flops, "ms_per_launch": cols_ms, "traffic": traffic,
            "traffic_source": traffic_src,
            "rows_kernel": {"achieved": (fl_alg["rows"] / kern["rows"]["launches"]) /
                            (kern["rows"]["ms_total"] / kern["rows"]["launches"] * 1e-3) / 1e12,
                            "ms_per_launch": kern["rows"]["ms_total"] / kern["rows"]["launches"]},
            "whole_step": {"achieved": step_ach * 1.0, "frac": step_ach / (best * world) if best else None,
                           "frac_of_nominal_peak": step_ach / (peak_nominal * world), "peak_is": f"{world} x the per-GPU peak",
                           "flops_per_image": algorithmic_flops(pn, N, n_src)["total"]},
            "hbm": {"compulsory_bytes_per_image": 8 * pn * pn * 2 + 8 * n_src + 4 * pn * pn,
                    "dram_bytes_per_image_measured": dram_per_image,
                    "dram_gbs_at_this_rate": (dram_per_image / (ms_per_step * 1e-3) / 1e9 / world) if dram_per_image else None,
                    "dram_note": "ncu dram__bytes of the row + column pass (profiles/traffic.json) scaled to the image's "
                                 "source points: the intermediate T (16.8 MB per source point) makes one HBM round trip",
                    "peak_gbs": hbm_peak, "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback"},
        }
        # our kernels launched by rank 0 inside the timed region (checked against profiles/r03u_launches.csv): per image
        # rows + cols per batch and, on the fast path with a coarse grid, rim_kernel + rim_reduce_kernel; per image this
        # rank post-processes (every N-th): coarse_unperm, 2 row passes, 3 column launches, assemble, finalize (7; 1 on
        # the generic path / without a coarse grid) and, with N > 1 and the peer sum, peer_sum_kernel
        coarse = plan.path == 2 and N // (2 * plan.M) > 1
        Sr_, Sc_ = plan.bbox[1] - plan.bbox[0] + 1, plan.bbox[3] - plan.bbox[2] + 1
        rim = 2 if (coarse and (Sr_ > plan.M or Sc_ > plan.M)) else 0
        per_image = 2 * ((n_mine + batch - 1) // batch) + rim
        per_root = (7 if coarse else 1) + (1 if pipe.reduce == "peer" else 0)
        rooted = len([i for i in range(args.steps) if i % world == 0])
        launches_total = per_image * args.steps + per_root * rooted
        cpu, library = None, None
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            from oracle import ref_runner as RR
            if RR.available():
                n_hi = max(8, args.ref_sample)
                vals, _, _, spent, (n_lo, n_hi), _ = time_reference(cfg, "cpu", max(2, n_hi // 4), n_hi, 1, 60.0, threads)
                cpu = {"value": float(np.median(vals)), "unit": UNIT, "cores": threads, "kind": "reference",
                       "sample": f"unmodified reference (oracle/_ref) abbeImage(fft=True, device=cpu), torch {threads} threads: "
                                 f"sub-sources of {n_lo} and {n_hi} of the {n_src} source points ({spent:.1f} s), image time = "
                                 "intercept + slope*n_src"}
            else:
                _, o_mft, o_pf, o_ls = build_inputs_host(args.config)   # the baseline leg builds its own inputs
                v, ns, _, dt = time_cpu_baseline(cfg, o_mft, o_pf, o_ls, max(threads, min(args.ref_sample, 4 * threads)),
                                                 threads)
                cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                       "sample": f"{ns} of {n_src} source points ({dt:.1f} s), extrapolated linearly in n_src; "
                                 "oracle numpy port, one source point per host thread (oracle/_ref not staged)"}
            # the Blackwell library baseline: the same unmodified reference handed
            # device='cuda' -- stock ATen kernels + cuFFT ifft2 on this very GPU, same inputs, after our timed region
            if RR.available() and not args.no_library_baseline:
                try:
                    vals, _, _, spent, (n_lo, n_hi), (ls_hi, ref_img) = time_reference(cfg, dev, 16, 80, 3, 40.0)
                    ours_sub = eng.abbe_fft(mft_d, pf_d, ls_hi.to(torch.int64), cfg.pixel_size, 4 / pn, cfg.wavelength)
                    library = {"value": float(np.median(vals)), "unit": UNIT, "what": "unmodified reference (oracle/_ref) "
                               "abbeImage(fft=True, device=cuda): stock ATen + cuFFT on the same B200, inputs from the "
                               "reference's own builders", "sample": f"sub-sources of {n_lo} and {n_hi} of the {n_src} source "
                               f"points, image time = intercept + slope*n_src, {len(vals)} repeats, {spent:.1f} s",
                               "rel_l2_product_vs_library_same_sub_source":
                                   float((ours_sub.double() - ref_img.double()).norm() / ref_img.double().norm())}
                except Exception as e:   # the baseline is a reported number, never a reason to lose the bench line
                    library = {"unavailable": f"{type(e).__name__}: {e}"[:300]}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": warm, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "complex64", "data": "synthetic",
                "config": {"workload": workload_string(cfg, n_src, N),
                           "l2": "flushed (160 MB write > 126 MB L2) between images, inside the timed region", "batch": batch,
                           "pipeline": "sequential, all-reduce, every rank post-processes" if args.no_pipeline else
                           "post-processing of image i overlaps the accumulation of image i+1 (2 streams); with N>1 "
                           "rank i mod N sums the partial planes and alone post-processes image i",
                           "subfft": plan.M, "residues": plan.R,
                           "path": "fast coarse-grid (2 FFTs of length M per line, spectral interpolation once per image)"
                           if plan.path == 2 else "generic fine-grid",
                           "sharding": f"source points interleaved over {world} rank(s)" + ("" if world == 1 else (
                               "; partial planes summed in rank order by the root with loads over NVLink from the peers' "
                               "CUDA-IPC-mapped planes (litho_peer_sum), sequence flags instead of a collective"
                               if pipe.reduce == "peer" else "; one ncclReduce of the intensity plane per image"))},
                "clocks": clocks, "gpu_launches": launches_total,
                "e2e": {"value": 1.0 / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "steps": e2e_steps, "rel_l2_vs_resident_path": e2e_check,
                        "how": "pinned host tensors -> AbbeEngine.prepare() (H2D + source-point extraction on a copy "
                               "stream, one image ahead" + ("; each rank uploads 1/N of every tensor, slices all-gathered "
                               "over NVLink (NCCL)" if upload_pg is not None else "") + ("; each rank uploads 1/N of the "
                               "input bytes, the other slices are pulled from the peers' CUDA-IPC-mapped staging buffers by "
                               "the copy engines (litho_peer_copy)" if upload_peers is not None else "") +
                               ") -> accumulate -> sum -> finalize "
                               "-> D2H on the post-processing stream; wall clock over the loop, max over ranks; "
                               "h2d_bytes_per_step is the whole job's"},
                "roofline": roofline, "cpu_baseline": cpu, "library_baseline": library, "parity": parity,
                "breakdown_ms": breakdown,
                "wall_s_timed_region": t_wall}
        if trace is not None:
            line["trace"] = trace
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="cfg3")
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--ref-sample", type=int, default=32)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-library-baseline", action="store_true")
    ap.add_argument("--generic", action="store_true", help="force the generic fine-grid kernels")
    ap.add_argument("--upload", default="peer", choices=["peer", "nccl", "full"],
                    help="e2e, N>1: each rank uploads 1/N of the inputs and the slices are exchanged by copy-engine "
                         "peer copies (default) or an NCCL all-gather; full: every rank uploads everything over PCIe")
    ap.add_argument("--no-chain", action="store_true", help="row pass of image i+1 waits for image i's last column pass")
    ap.add_argument("--no-pipeline", action="store_true", help="finalize each image before starting the next")
    ap.add_argument("--trace", action="store_true", help="extra untimed loop with per-image CUDA events on every rank")
    ap.add_argument("--reduce", default="peer", choices=["peer", "nccl"],
                    help="N>1: sum of the partial planes by peer-memory loads on the root (default) or ncclReduce")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
