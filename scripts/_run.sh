nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r01y_gpu_tests.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/r01y_gpu_tests.log
scripts/gpu_ab.sh r01y "LITHO_TMA=0|" "LITHO_TMA=1|" "LITHO_TMA=1|--batch 8" "LITHO_TMA=1|--batch 4" "LITHO_TMA=1 LITHO_TMA_L2=256|" "LITHO_TMA=1 LITHO_TMA_L2=64|"
