mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q -k "two_gpu" > gpurun_out/r02d_tests.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/r02d_tests.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 16 --warmup 8 > gpurun_out/r02d_n2.log 2>&1; tail -1 gpurun_out/r02d_n2.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e'])"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 16 --warmup 8 --full-upload > gpurun_out/r02d_n2_full.log 2>&1; tail -1 gpurun_out/r02d_n2_full.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e'])"
