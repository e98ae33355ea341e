timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r02i_gpu_tests.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/r02i_gpu_tests.log
timeout 900 python scripts/sweep_cfg5.py --focus 2 > gpurun_out/r02i_cfg5_n1_f2.log 2>&1; tail -1 gpurun_out/r02i_cfg5_n1_f2.log | cut -c1-500
scripts/gpu_ab.sh r02i "X=1|"
