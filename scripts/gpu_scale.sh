#!/bin/bash
# Scaling curve on ONE 8-GPU box (gpurun --gpus 8): bench.py at N = 1, 2, 4, 8 with the driver's flags, cfg4 at 8 and 1,
# the cfg5 focus sweep at 8.  Usage: scripts/gpu_scale.sh TAG
TAG=${1:-scale}
mkdir -p gpurun_out
run() {  # name, nproc, extra flags
  if [ "$2" = "1" ]; then
    timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-library-baseline $3 > gpurun_out/${TAG}_$1.log 2>&1
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $2 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 200)) \
      bench.py --gpus $2 --steps 20 --warmup 5 --no-cpu-baseline $3 > gpurun_out/${TAG}_$1.log 2>&1
  fi
  echo "== $1 exit $?"; grep '^{"metric"' gpurun_out/${TAG}_$1.log | tail -1 | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read())
    print({k: d.get(k) for k in ('value', 'ms_per_step', 'n_gpus')}, 'e2e', d['e2e']['value'], 'parity', d.get('parity', {}).get('rel_l2_vs_reference_golden'), 'frac', d['roofline']['whole_step']['frac'])
except Exception as e:
    print('unparsed', e)
"
}
run n1 1
run n2 2
run n4 4
run n8_a 8
run n8_b 8
run n8_nccl 8 "--reduce nccl"
run cfg4_n8 8 "--config cfg4 --steps 5 --warmup 3"
run cfg4_n1 1 "--config cfg4 --steps 3 --warmup 3"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29871 \
    scripts/sweep_cfg5.py --repeats 2 > gpurun_out/${TAG}_cfg5_sweep_n8.log 2>&1
echo "== cfg5 sweep exit $?"; grep '^{"workload"' gpurun_out/${TAG}_cfg5_sweep_n8.log | tail -1
