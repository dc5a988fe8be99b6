cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
{ timeout 300 python -m pytest tests/test_gpu_parity_r2.py -x -q -m gpu -k "groupwise" 2>&1 | tail -3;
timeout 300 python tools/gs_tune.py 2>&1 | cut -c1-160; } > gpurun_out/r2_gsB.log 2>&1
