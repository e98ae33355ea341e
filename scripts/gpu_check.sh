#!/bin/bash
# One-GPU check on the B200 box via gpurun: parity tests, smoke, default bench (+ optional ncu launch list and full
# capture of the two hot kernels).  Usage: scripts/gpu_check.sh [tag]   (SKIP_NCU=1 to leave the profiler out)
TAG=${1:-r03}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_gpu_tests.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/${TAG}_smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench.log 2>&1
echo "bench exit $?" >> gpurun_out/${TAG}_bench.log
if [ "${SKIP_NCU:-0}" != "1" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 400 --csv \
    --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:abbe_fast -s 8 -c 4 \
    -o gpurun_out/${TAG}_prof python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_ncu_full.log 2>&1
fi
tail -15 gpurun_out/${TAG}_gpu_tests.log | cut -c1-300; cat gpurun_out/${TAG}_smoke.log | tail -2; tail -c 5000 gpurun_out/${TAG}_bench.log
if [ "${OTHER_CONFIGS:-0}" = "1" ]; then
for c in cfg4 cfg5; do
  timeout 600 python bench.py --config $c --steps 3 --warmup 3 --no-cpu-baseline --no-library-baseline > gpurun_out/${TAG}_bench_$c.log 2>&1
  echo "== $c"; tail -1 gpurun_out/${TAG}_bench_$c.log | cut -c1-600
done
fi
