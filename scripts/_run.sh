timeout 600 python -m pytest tests -m gpu -x -q -k "tma or pipeline or golden or cfg3" > gpurun_out/r02b_gpu_tests.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/r02b_gpu_tests.log
scripts/gpu_ab.sh r02b "LITHO_COL_SPLIT=1|" "LITHO_COL_SPLIT=2|" "LITHO_COL_SPLIT=2|--batch 12"
