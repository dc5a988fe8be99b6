#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 120 python tools/ncu_dec.py 1 4096 11008 > gpurun_out/r2_sanity7.log 2>&1; echo "sanity rc=$?" >> gpurun_out/r2_sanity7.log
timeout 900 python -m pytest tests/test_gpu_linear.py tests/test_gpu_parity_r2.py -x -q -m gpu > gpurun_out/r2_pytest7.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_pytest7.log
timeout 600 python tools/dec_tune.py '[[1,4096,11008,{}],[1,4096,11008,{"gemv_pf":0}],[1,4096,11008,{"gemv_pf":1}],[1,4096,11008,{"gemv_early":1}],[1,4096,11008,{"gemv_early":3}],[1,4096,11008,{"gemv_early":5}],[2,4096,11008,{}],[3,4096,11008,{}],[4,4096,11008,{}],[8,4096,11008,{}],[16,4096,11008,{}],[1,11008,4096,{}],[1,11008,4096,{"gemv_early":1}],[1,11008,4096,{"gemv_early":3}],[2,11008,4096,{}],[4,11008,4096,{}],[8,11008,4096,{}],[16,11008,4096,{}]]' > gpurun_out/r2_tune7.jsonl 2> gpurun_out/r2_tune7.err
export B200Q_LIB=$PWD/fused-4-bit-dequantize-linear-cuda-kernel_b200/libb200q_prof.so
for cfg in "1 4096 11008" "4 4096 11008" "8 4096 11008"; do
  echo "== $cfg"; timeout 120 python tools/prof_dec.py $cfg 2>&1 | tail -18
done > gpurun_out/r2_prof7.log 2>&1
