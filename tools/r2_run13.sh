#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ep.py tests/test_gpu_moe.py -x -q -m gpu > gpurun_out/r2_pytest13.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_pytest13.log
for N in 8 4 2; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2964$N bench.py --gpus $N --steps 10 --warmup 3 --no-cpu > gpurun_out/r2_bench13_n$N.json 2> gpurun_out/r2_bench13_n$N.err
echo "bench rc=$?" >> gpurun_out/r2_bench13_n$N.err
done
