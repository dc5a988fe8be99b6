"""Recipe: stage the REAL reference next to the oracle.                       *** TEST INFRASTRUCTURE ***

    python oracle/make_ref.py            (also run by __graft_entry__.build() when /root/reference is present)

  oracle/_ref/reference/   the reference's pure-Python path, copied from where it lies under /root/reference
                           (python/, tests/, benchmark/moe_grouped_gemm/).  Git-ignored -- reference sources never
                           enter this repo's history -- but not gpurun-ignored, so it travels to the GPU box, where
                           /root/reference does not exist.  Used by
                             * tests/test_gpu_reference_suite.py: the reference's own tests/test_correctness.py run
                               UNMODIFIED against this repo's drop-in `fused_quant_linear_cuda` shim;
                             * tests/test_oracle.py: the oracle restatement checked against the live reference;
                             * bench.py --impl reference / cpu_baseline: the reference's CPU path itself
                               (python/quantize.py dequantize_weights + F.linear), "kind": "reference".
  baseline/_ref/           the reference's CUDA extensions built for sm_100 with its own setup.py
                           (pip install --no-index --target, from a scratch copy because the build writes into the
                           tree): the same-box "reference GPU kernel" row of bench.py (ref_gpu_kernel).

Nothing under either directory is imported by the product package.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.environ.get("B200Q_REFERENCE", "/root/reference")
DST = os.path.join(HERE, "_ref", "reference")
PARTS = ["python", "tests", os.path.join("benchmark", "moe_grouped_gemm")]


def stage_python():
    if not os.path.isdir(SRC):
        print(f"make_ref: {SRC} not present, keeping whatever oracle/_ref holds")
        return False
    for part in PARTS:
        dst = os.path.join(DST, part)
        shutil.rmtree(dst, ignore_errors=True)
        shutil.copytree(os.path.join(SRC, part), dst, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    # benchmark/ has no __init__.py in the reference either: `benchmark.moe_grouped_gemm` is a namespace import
    with open(os.path.join(HERE, "_ref", "PROVENANCE.txt"), "w") as f:
        f.write(f"copied by oracle/make_ref.py from {SRC} ({', '.join(PARTS)}); not tracked by git\n")
    return True


def build_cuda_ext(force=False):
    """The reference's two CUDAExtensions for sm_100 -> baseline/_ref/*.so (minutes: torch/extension.h)."""
    out = os.path.join(ROOT, "baseline", "_ref")
    if not force and os.path.isdir(out) and any(n.startswith("fused_quant_linear_cuda") and n.endswith(".so") for n in os.listdir(out)):
        return True
    if not os.path.isdir(SRC):
        return False
    tmp = "/tmp/b200q_refsrc"
    shutil.rmtree(tmp, ignore_errors=True)
    shutil.copytree(SRC, tmp, ignore=shutil.ignore_patterns("__pycache__"))
    env = dict(os.environ, TORCH_CUDA_ARCH_LIST="10.0", MAX_JOBS="4")
    cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps",
           "--find-links", "/opt/wheelhouse", "--target", out, "."]
    r = subprocess.run(cmd, cwd=tmp, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, "INSTALL.log"), "w") as f:
        f.write(" ".join(cmd) + "\n" + r.stdout[-4000:])
    return r.returncode == 0


if __name__ == "__main__":
    ok = stage_python()
    print("reference python path staged:", ok, "->", DST)
    if "--cuda" in sys.argv:
        print("reference CUDA extensions built:", build_cuda_ext(force="--force" in sys.argv))
