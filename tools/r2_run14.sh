#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_linear.py tests/test_gpu_parity_r2.py -x -q -m gpu > gpurun_out/r2_pytest14.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_pytest14.log
timeout 600 python tools/dec_tune.py '[[1,4096,11008,{}],[1,4096,11008,{"gemv_pf":0}],[1,4096,11008,{"gemv_early":93}],[1,4096,11008,{"gemv_early":91}],[2,4096,11008,{}],[4,4096,11008,{}],[1,11008,4096,{}],[2,11008,4096,{}],[1,4096,11008,{}]]' > gpurun_out/r2_tune14.jsonl 2> gpurun_out/r2_tune14.err
