cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
HINT=1 timeout 120 python tools/ncu_dec.py 8 4096 11008 > gpurun_out/r02_ncu_hm.plain.log 2>&1 || exit 1
HINT=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemv_hm -s 24 -c 4 -o gpurun_out/r02_gemv_hm_m8 -f python tools/ncu_dec.py 8 4096 11008 > gpurun_out/r02_ncu_hm.log 2>&1
echo "rc=$?" >> gpurun_out/r02_ncu_hm.log
HINT=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemv_hm -s 24 -c 4 -o gpurun_out/r02_gemv_hm_m4 -f python tools/ncu_dec.py 4 4096 11008 > gpurun_out/r02_ncu_hm4.log 2>&1
