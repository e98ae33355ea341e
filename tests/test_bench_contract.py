"""The reference arm of bench.py (CPU only): `bench.py --impl reference` must run the UNMODIFIED reference staged under
oracle/_ref (or the numpy port where it is not staged), print ONE JSON line with the contract's keys, and use the same
workload string as the GPU arm.  Uses the small cfg1 so that it takes seconds."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_line_cfg1():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "cfg1",
                          "--steps", "2", "--warmup", "1", "--ref-sample", "16"], capture_output=True, text=True,
                         timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "impl", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["value"] > 0
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"]
    sys.path.insert(0, ROOT)
    import bench
    from lithographysimulator_b200 import workloads as wl
    from oracle import build_ref
    cfg = wl.CONFIGS["cfg1"]
    assert d["config"]["workload"] == bench.workload_string(cfg, 92, 512)      # the GPU arm builds the same string
    if build_ref.staged():
        assert cb["kind"] == "reference"


def test_ranks_other_than_zero_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "cfg1", "--gpus", "2"],
                         capture_output=True, text=True, timeout=300, cwd=ROOT, env=env)
    assert out.returncode == 0 and not [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
