#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
export B200Q_LIB=$PWD/fused-4-bit-dequantize-linear-cuda-kernel_b200/libb200q_prof.so
export B200Q_TUNE=gemv_tma3d=1
for dbg in 1 3 5 7 9 17 19 31; do
  echo "== M=1 DBG=$dbg"; DBG=$dbg timeout 120 python tools/prof_dec.py 1 4096 11008 2>&1 | tail -17
done > gpurun_out/r2_prof2.log 2>&1
