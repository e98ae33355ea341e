#!/bin/bash
# First GPU call of round 2 (one B200): the checks prepared at the end of round 1.
#   1. full cfg4 / cfg5 images against the reference's (tests/golden/cfg4.npz, cfg5.npz)
#   2. compute-sanitizer memcheck/racecheck/synccheck on the small parity cases
#   3. default bench (now with library_baseline = the unmodified reference on device='cuda', parity vs golden)
#   4. f2 evidence: direct solver FP32 vs tensor-core GEMM proxies
# Usage: gpurun --timeout 2400 -- 'bash scripts/gpu_next.sh r03a'
TAG=${1:-r03a}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_gpu_tests.log 2>&1
echo "pytest exit $?" | tee -a gpurun_out/${TAG}_gpu_tests.log; tail -4 gpurun_out/${TAG}_gpu_tests.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench.log 2>&1
echo "bench exit $?"; tail -c 4000 gpurun_out/${TAG}_bench.log
timeout 400 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${TAG}_bench_ref.log 2>&1
echo "bench ref exit $?"; tail -c 1500 gpurun_out/${TAG}_bench_ref.log
timeout 300 python scripts/direct_bench.py > gpurun_out/${TAG}_direct.log 2>&1
echo "direct exit $?"; tail -c 3000 gpurun_out/${TAG}_direct.log
timeout 900 bash scripts/sanitize.sh gpu > gpurun_out/${TAG}_sanitizer.log 2>&1
echo "compute-sanitizer: exit $?"; tail -5 gpurun_out/${TAG}_sanitizer.log
