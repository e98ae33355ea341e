#!/bin/bash
# f2 evidence under ncu: the tcgen05 direct-solver kernel and the FP32 CUDA-core kernels it replaces, at pn = 256.
TAG=${1:-direct}
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:direct_tc_kernel -s 2 -c 4 \
    -o gpurun_out/${TAG}_tc python scripts/direct_bench.py --pn 256 > gpurun_out/${TAG}_tc.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:direct_(rows|cols)_kernel" -s 2 -c 4 \
    -o gpurun_out/${TAG}_fp32 python scripts/direct_bench.py --pn 256 > gpurun_out/${TAG}_fp32.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:direct|gemm|cutlass|sm100|tensorop" -c 200 --csv \
    --log-file gpurun_out/${TAG}_launches.csv python scripts/direct_bench.py --pn 256 > gpurun_out/${TAG}_launch.log 2>&1
timeout 300 python scripts/direct_bench.py > gpurun_out/${TAG}_bench.log 2>&1
tail -1 gpurun_out/${TAG}_bench.log | cut -c1-200
ls -la gpurun_out/${TAG}_*
