#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2_gpus9.log 2>&1
timeout 900 python -m pytest tests/test_gpu_ep.py -x -q -m gpu > gpurun_out/r2_pytest9.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_pytest9.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_bench9_n2.json 2> gpurun_out/r2_bench9_n2.err
echo "bench rc=$?" >> gpurun_out/r2_bench9_n2.err
