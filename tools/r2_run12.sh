#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
N=${1:-8}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29631 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu > gpurun_out/r2_bench12_n$N.json 2> gpurun_out/r2_bench12_n$N.err
echo "bench rc=$?" >> gpurun_out/r2_bench12_n$N.err
