#!/bin/bash
# batch-size sweep of the headline bench (no ncu): prints images/s per batch
TAG=${1:-sweep}
mkdir -p gpurun_out
for b in ${BATCHES:-1 2 3 4 6 8}; do
  timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --batch $b > gpurun_out/${TAG}_b$b.log 2>&1
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_b$b.log").read().strip().splitlines()[-1])
    r=d["roofline"]
    print("batch $b: %.2f img/s  %.2f ms  e2e %.2f  cols %.1f us/launch rows %.1f us/launch" % (d["value"], d["ms_per_step"], d["e2e"]["value"], r["ms_per_launch"]*1e3, r["rows_kernel"]["ms_per_launch"]*1e3))
except Exception as e:
    print("batch $b failed", e); print(open("gpurun_out/${TAG}_b$b.log").read()[-800:])
PY
done
