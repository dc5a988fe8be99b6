#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python tools/dec_tune.py '[[1,4096,11008,{"gemv_pf":0}],[1,4096,11008,{"gemv_pf":3}],[1,4096,11008,{"gemv_pf":3,"gemv_early":1}],[1,4096,11008,{"gemv_pf":3,"gemv_early":3}],[1,4096,11008,{"gemv_pf":3,"gemv_early":5}],[1,4096,11008,{"gemv_pf":3,"gemv_early":2}],[1,4096,11008,{"gemv_pf":3,"gemv_early":93}],[1,4096,11008,{"gemv_pf":3,"gemv_early":91}],[1,4096,11008,{"gemv_pf":0,"gemv_early":5}],[1,4096,11008,{"gemv_pf":0,"gemv_early":93}],[2,4096,11008,{"gemv_pf":3}],[4,4096,11008,{"gemv_pf":3}],[8,4096,11008,{"gemv_pf":3}],[16,4096,11008,{"gemv_pf":3}],[1,11008,4096,{"gemv_pf":3}],[1,11008,4096,{"gemv_pf":0}],[1,11008,4096,{"gemv_pf":3,"gemv_early":1}],[1,11008,4096,{"gemv_pf":3,"gemv_early":91}],[2,11008,4096,{"gemv_pf":3}],[4,11008,4096,{"gemv_pf":3}]]' > gpurun_out/r2_tune2.jsonl 2> gpurun_out/r2_tune2.err
export B200Q_LIB=$PWD/fused-4-bit-dequantize-linear-cuda-kernel_b200/libb200q_prof.so
export B200Q_TUNE=gemv_pf=3
for cfg in "1 4096 11008" "1 11008 4096"; do
  echo "== $cfg"; timeout 120 python tools/prof_dec.py $cfg 2>&1 | tail -18
done > gpurun_out/r2_prof4.log 2>&1
