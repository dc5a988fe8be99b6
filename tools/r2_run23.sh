#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
export B200Q_LIB=$PWD/fused-4-bit-dequantize-linear-cuda-kernel_b200/libb200q_prof.so
for cfg in "1 4096 11008" "2 4096 11008" "4 4096 11008" "1 11008 4096"; do
  echo "=== $cfg" >> gpurun_out/r2_prof23.log
  DBG=1 timeout 120 python tools/prof_dec.py $cfg >> gpurun_out/r2_prof23.log 2>&1
done
