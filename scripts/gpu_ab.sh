#!/bin/bash
# A/B on ONE box: the previous build (.ab_old/, git-ignored copy of an older commit with its own .so) against the
# working tree, alternating, same bench command.  Usage: gpurun -- 'bash scripts/gpu_ab.sh tag'
TAG=${1:-ab}
mkdir -p gpurun_out
for rep in 1 2; do
  (cd .ab_old && timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline) > gpurun_out/${TAG}_old_$rep.log 2>&1
  timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-library-baseline --trace > gpurun_out/${TAG}_new_$rep.log 2>&1
done
for f in gpurun_out/${TAG}_old_1.log gpurun_out/${TAG}_new_1.log gpurun_out/${TAG}_old_2.log gpurun_out/${TAG}_new_2.log; do
  echo "== $f"; tail -1 $f | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read())
    print({k: d.get(k) for k in ('value', 'ms_per_step')}, d['e2e']['value'], d['breakdown_ms'], d['roofline']['ms_per_launch'], d['roofline']['rows_kernel']['ms_per_launch'], d.get('trace'))
except Exception as e:
    print('unparsed', e)
"
done
