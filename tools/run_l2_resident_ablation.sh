#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
CFG='[[1,4096,11008,{}],[1,4096,11008,{"gemv_pf":0}],[2,4096,11008,{}],[1,11008,4096,{}]]'
echo "pool 24" > gpurun_out/r2_l2test.log; python tools/dec_tune.py "$CFG" >> gpurun_out/r2_l2test.log 2>&1
echo "pool 2 (L2 resident)" >> gpurun_out/r2_l2test.log; POOL_LAYERS=2 python tools/dec_tune.py "$CFG" >> gpurun_out/r2_l2test.log 2>&1
echo "pool 4 (L2 resident)" >> gpurun_out/r2_l2test.log; POOL_LAYERS=4 python tools/dec_tune.py "$CFG" >> gpurun_out/r2_l2test.log 2>&1
export B200Q_LIB=$PWD/fused-4-bit-dequantize-linear-cuda-kernel_b200/libb200q_prof.so
echo "prof build, normal" >> gpurun_out/r2_l2test.log; DBG=1 python tools/prof_dec.py 1 4096 11008 2>&1 | grep us_per >> gpurun_out/r2_l2test.log
echo "prof build, no weight traffic" >> gpurun_out/r2_l2test.log; DBG=5 python tools/prof_dec.py 1 4096 11008 2>&1 | tail -16 >> gpurun_out/r2_l2test.log
