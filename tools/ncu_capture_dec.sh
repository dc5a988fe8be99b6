#!/bin/bash
# one ncu --set full capture per call: $1 = hint (1/0), $2 = tag
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
HINT=$1 timeout 120 python tools/ncu_dec.py 1 4096 11008 > gpurun_out/r02_ncu_dec_$2.plain.log 2>&1 || exit 1
HINT=$1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemv_dec -s 24 -c 12 -o gpurun_out/r02_gemv_dec_$2 -f python tools/ncu_dec.py 1 4096 11008 > gpurun_out/r02_ncu_dec_$2.log 2>&1
echo "rc=$?" >> gpurun_out/r02_ncu_dec_$2.log
