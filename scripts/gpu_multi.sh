#!/bin/bash
# Multi-GPU check on one box (gpurun --gpus N): the 2-GPU parity test, then bench.py at N GPUs with the peer-memory
# sum (default) and with ncclReduce, driver-style flags, repeated.  Usage: scripts/gpu_multi.sh TAG N [REPS] [CONFIG]
TAG=${1:-multi}; N=${2:-2}; REPS=${3:-2}; CFG=${4:-cfg3}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,clocks.sm --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
nvidia-smi topo -m >> gpurun_out/${TAG}_gpu.txt 2>&1
if [ "${SKIP_TEST:-0}" != "1" ]; then
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "two_gpu" > gpurun_out/${TAG}_tests.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_tests.log; tail -15 gpurun_out/${TAG}_tests.log | cut -c1-300
fi
run() {  # name, extra flags, env assignments
  timeout 600 env $3 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 200)) \
      bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline --config $CFG $2 > gpurun_out/${TAG}_$1.log 2>&1
  echo "== $1 exit $?"; grep '^{"metric"' gpurun_out/${TAG}_$1.log | tail -1 | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read())
    print({k: d.get(k) for k in ('value', 'ms_per_step', 'n_gpus')}, 'e2e', d['e2e']['value'], d['breakdown_ms'], 'parity', d.get('parity', {}).get('rel_l2_vs_reference_golden'), d.get('trace'))
except Exception as e:
    print('unparsed', e)
"
}
for rep in $(seq 1 $REPS); do
  run peer_$rep "--trace"
  run nccl_$rep "--reduce nccl --trace"
done
if [ "${VARIANTS:-1}" = "1" ]; then
  run peer_prio0_slots2 "--trace" "LITHO_FIN_PRIORITY=0 LITHO_PLANE_SLOTS=2"
  run peer_steps5 "--steps 5 --warmup 3"
fi
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-library-baseline --config $CFG > gpurun_out/${TAG}_n1.log 2>&1
echo "== n1"; tail -1 gpurun_out/${TAG}_n1.log | cut -c1-200
