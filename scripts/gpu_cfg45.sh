#!/bin/bash
# cfg4 / cfg5 on one GPU: parity tests of the big grids, then bench lines with the TMA-staged column kernel on and off.
TAG=${1:-cfg45}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "cfg5 or cfg4 or largest" > gpurun_out/${TAG}_tests.log 2>&1
echo "pytest exit $?"; tail -4 gpurun_out/${TAG}_tests.log | cut -c1-300
for c in cfg5 cfg4; do
  for tma in 1 0; do
    LITHO_TMA=$tma timeout 600 python bench.py --config $c --steps 4 --warmup 3 --no-cpu-baseline --no-library-baseline > gpurun_out/${TAG}_${c}_tma$tma.log 2>&1
    echo "== $c tma=$tma"; tail -1 gpurun_out/${TAG}_${c}_tma$tma.log | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read())
    print({k: d.get(k) for k in ('value', 'ms_per_step')}, 'e2e', d['e2e']['value'], 'cols', d['roofline']['ms_per_launch'], 'frac', d['roofline']['frac'], 'rows', d['roofline']['rows_kernel']['ms_per_launch'], 'step frac', d['roofline']['whole_step']['frac'], 'parity', d.get('parity', {}) and d['parity'].get('rel_l2_vs_reference_golden'), 'batch', d['config']['batch'])
except Exception as e:
    print('unparsed', e)
"
  done
done
