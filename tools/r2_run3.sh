#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_linear.py tests/test_gpu_parity_r2.py -x -q -m gpu > gpurun_out/r2_pytest3.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_pytest3.log
timeout 600 python tools/dec_tune.py > gpurun_out/r2_tune1.jsonl 2> gpurun_out/r2_tune1.err
