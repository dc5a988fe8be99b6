#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_linear.py tests/test_gpu_parity_r2.py -x -q -m gpu > gpurun_out/r2_pytest2.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_pytest2.log
timeout 300 python tools/sweep_all.py decode > gpurun_out/r2_sweep_v2.jsonl 2> gpurun_out/r2_sweep_v2.err
export B200Q_LIB=$PWD/fused-4-bit-dequantize-linear-cuda-kernel_b200/libb200q_prof.so
for cfg in "1 4096 11008" "4 4096 11008" "8 4096 11008" "1 11008 4096"; do
  echo "== $cfg"; timeout 120 python tools/prof_dec.py $cfg 2>&1 | tail -18
done > gpurun_out/r2_prof3.log 2>&1
