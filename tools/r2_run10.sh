#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu > gpurun_out/r2_bench10_n8.json 2> gpurun_out/r2_bench10_n8.err
echo "bench rc=$?" >> gpurun_out/r2_bench10_n8.err
