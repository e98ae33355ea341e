#!/bin/bash
# A/B runs of the headline bench under environment toggles (no ncu).  Usage: scripts/gpu_ab.sh TAG "ENV1=.. ENV2=..|--flags" ...
# Each argument after TAG is "<env assignments>|<bench flags>"; prints one summary line per run.
TAG=${1:-ab}; shift
mkdir -p gpurun_out
i=0
for spec in "$@"; do
  envs="${spec%%|*}"; flags="${spec#*|}"
  [ "$flags" == "$spec" ] && flags=""
  i=$((i+1))
  log=gpurun_out/${TAG}_$i.log
  env $envs timeout 400 python bench.py --steps ${STEPS:-8} --warmup 4 --no-cpu-baseline $flags > $log 2>&1
  python - "$log" "$spec" <<'PY'
import json, sys
log, spec = sys.argv[1], sys.argv[2]
try:
    d = json.loads(open(log).read().strip().splitlines()[-1])
    r = d["roofline"]
    print("%-40s %.2f img/s %.3f ms  e2e %.2f (chk %s) cols %.1f us rows %.1f us frac %.3f step-frac %.3f" % (
        spec, d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"].get("rel_l2_vs_resident_path"),
        r["ms_per_launch"] * 1e3, r["rows_kernel"]["ms_per_launch"] * 1e3, r["frac"], r["whole_step"]["frac"]))
except Exception as e:
    print(spec, "FAILED", e); print(open(log).read()[-1500:])
PY
done
