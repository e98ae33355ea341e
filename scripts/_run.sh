mkdir -p gpurun_out
python bench.py > gpurun_out/r02l_bench_default.log 2>&1; tail -1 gpurun_out/r02l_bench_default.log | cut -c1-200
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 300 --csv --log-file gpurun_out/r02l_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r02l_ncu_launch.log 2>&1
timeout 400 ncu --set full --cache-control none --clock-control none --import-source on -k regex:abbe_fast -s 12 -c 4 -o gpurun_out/r02l_prof python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r02l_prof.log 2>&1
ls -la gpurun_out | grep r02l
