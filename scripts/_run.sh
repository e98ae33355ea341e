mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
$T --master-port 29511 bench.py --gpus 8 --steps 16 --warmup 8 > gpurun_out/r02h_n8.log 2>&1; tail -1 gpurun_out/r02h_n8.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('cfg3 n8', d['value'], d['ms_per_step'], d['e2e']['value'], d['config']['batch'])"
$T --master-port 29512 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r02h_n8_k5.log 2>&1; tail -1 gpurun_out/r02h_n8_k5.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('cfg3 n8 k5', d['value'], d['ms_per_step'], d['e2e']['value'])"
$T --master-port 29513 bench.py --gpus 8 --steps 3 --warmup 3 --config cfg4 > gpurun_out/r02h_n8_cfg4.log 2>&1; tail -1 gpurun_out/r02h_n8_cfg4.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('cfg4 n8', d['value'], d['ms_per_step'], d['e2e']['value'])"
timeout 600 $T --master-port 29514 scripts/sweep_cfg5.py > gpurun_out/r02h_cfg5_n8.log 2>&1; tail -2 gpurun_out/r02h_cfg5_n8.log | cut -c1-600
python bench.py --steps 16 --warmup 8 > gpurun_out/r02h_n1.log 2>&1; tail -1 gpurun_out/r02h_n1.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('cfg3 n1', d['value'], d['ms_per_step'], d['e2e']['value'])"
python bench.py --steps 5 --warmup 3 > gpurun_out/r02h_n1_k5.log 2>&1; tail -1 gpurun_out/r02h_n1_k5.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('cfg3 n1 k5', d['value'], d['ms_per_step'], d['e2e']['value'])"
