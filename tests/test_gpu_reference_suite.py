"""The reference's OWN test files, run unmodified against this repo's drop-in surface.

oracle/_ref/reference/ holds a staged copy of the reference's python/ and tests/ (oracle/make_ref.py; git-ignored,
travels to the GPU box).  tests/test_correctness.py there does `from python.quantize import ...` (the reference's CPU
oracle) and, in TestCUDAKernel (tests/test_correctness.py:189-253), `import fused_quant_linear_cuda` -- which resolves
to the shim at this repo's root, i.e. to libb200q's kernels through the C ABI.  Nothing is patched: the three CUDA
tests compare our GPU results with the reference's dequantize + F.linear at the reference's tolerances."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "reference")

pytestmark = pytest.mark.gpu


def _run(args):
    if not os.path.isdir(os.path.join(REF, "tests")):
        pytest.skip("oracle/_ref/reference not staged (run `python oracle/make_ref.py` where /root/reference exists)")
    env = dict(os.environ)
    # the reference's tests insert their own root (REF) at sys.path[0]; the repo root supplies the extension shims
    env["PYTHONPATH"] = os.pathsep.join([ROOT, env.get("PYTHONPATH", "")])
    return subprocess.run([sys.executable, "-m", "pytest", "-q", "-p", "no:cacheprovider", "-rs"] + args, cwd=REF, env=env,
                          stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)


def test_reference_cuda_tests_run_unmodified_against_the_shim():
    r = _run(["tests/test_correctness.py", "-k", "TestCUDAKernel"])
    print(r.stdout[-3000:])
    assert r.returncode == 0, r.stdout[-3000:]
    assert "3 passed" in r.stdout and "skipped" not in r.stdout.splitlines()[-1], r.stdout[-1500:]


def test_reference_whole_suite_passes_on_the_box():
    """All 24 reference tests (21 CPU tests of its own oracle + the 3 CUDA tests through the shim)."""
    r = _run(["tests"])
    print(r.stdout[-3000:])
    assert r.returncode == 0 and "24 passed" in r.stdout, r.stdout[-3000:]
