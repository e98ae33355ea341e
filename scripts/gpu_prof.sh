#!/bin/bash
# ncu full capture of the fast kernels with caches left warm (closer to the live run than the default flush)
TAG=${1:-prof}
mkdir -p gpurun_out
timeout 900 ncu --set full --cache-control none --clock-control none --import-source on -k regex:abbe_fast -s ${SKIP:-20} -c ${COUNT:-4} \
    -o gpurun_out/${TAG} python bench.py --steps 1 --warmup 1 --no-cpu-baseline --batch ${BATCH:-12} > gpurun_out/${TAG}.log 2>&1
tail -3 gpurun_out/${TAG}.log | cut -c1-400
