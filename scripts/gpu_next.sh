#!/bin/bash
# First GPU call of the next round (one B200): the checks that were prepared at the end of round 1 without GPU time left.
#   1. full cfg4 / cfg5 images against the reference's (tests/golden/cfg4.npz, cfg5.npz) -- then drop the LITHO_FULL_GOLDEN gate
#   2. compute-sanitizer memcheck/racecheck/synccheck on the small parity cases
#   3. the stock torch/cuFFT library baseline of the hot loop next to the product
# Usage: gpurun --timeout 1500 -- 'bash scripts/gpu_next.sh'
mkdir -p gpurun_out
LITHO_FULL_GOLDEN=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "full_cfg4_cfg5" > gpurun_out/next_full_golden.log 2>&1
echo "full goldens: exit $?"; tail -3 gpurun_out/next_full_golden.log
timeout 900 bash scripts/sanitize.sh gpu > gpurun_out/next_sanitizer.log 2>&1
echo "compute-sanitizer: exit $?"; tail -3 gpurun_out/next_sanitizer.log
timeout 300 python scripts/torch_baseline.py --config cfg3 --points 64 > gpurun_out/next_torch_baseline.log 2>&1
tail -1 gpurun_out/next_torch_baseline.log
