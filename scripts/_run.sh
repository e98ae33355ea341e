mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
$T --master-port 29511 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r02k_n8_k5.log 2>&1; tail -1 gpurun_out/r02k_n8_k5.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('cfg3 n8 k5', d['value'], d['ms_per_step'], d['e2e']['value'])"
