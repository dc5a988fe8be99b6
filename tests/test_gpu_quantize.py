"""GPU parity for the quantisation format kernels: packing, scales and zero points bit-exact with
the reference (python/quantize.py:38-124), dequantisation bit-exact (python/quantize.py:127-173)."""
import hashlib

import numpy as np
import pytest
import torch

from conftest import load_golden

pytestmark = pytest.mark.gpu
Q = load_golden("quantize")
MI = load_golden("moe_int4_module")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.mark.parametrize("name", ["s42_16x32", "s123_256x512", "s42_64x128", "s42_256x512", "edge"])
def test_quantize_golden_bit_exact(pkg, name):
    w = torch.from_numpy(Q[f"{name}_w"]).cuda()
    p, s, z = pkg.quantize_weights(w)
    assert p.shape == (w.shape[0], w.shape[1] // 2) and p.dtype == torch.uint8
    assert np.array_equal(p.cpu().numpy(), Q[f"{name}_packed"])
    assert np.array_equal(s.cpu().numpy().view(np.uint32), Q[f"{name}_scales"].view(np.uint32))
    assert np.array_equal(z.cpu().numpy(), Q[f"{name}_zp"])
    d = pkg.dequantize_weights(p, s, z)
    assert np.array_equal(d.cpu().numpy().view(np.uint32), Q[f"{name}_deq"].view(np.uint32))


def test_quantize_4096_seed7_digest(pkg):
    torch.manual_seed(7)
    w = torch.randn(4096, 4096)
    p, s, z = pkg.quantize_weights(w.cuda())
    assert sha(p.cpu().numpy()) == str(Q["s7_4096_packed_sha"])
    assert np.array_equal(s.cpu().numpy(), Q["s7_4096_scales"])
    assert np.array_equal(z.cpu().numpy(), Q["s7_4096_zp"])
    # the reference's round-trip bound (tests/test_correctness.py:68-75)
    assert torch.allclose(w, pkg.dequantize_weights(p, s, z).cpu(), atol=0.5)


def test_from_linear_llama_up_digest(pkg):
    """BASELINE configs 1/2: QuantizedLinear.from_linear(nn.Linear(4096, 11008)), seed 42; host
    weights are quantised on the GPU and the buffers come back on the weights' device."""
    torch.manual_seed(42)
    lin = torch.nn.Linear(4096, 11008, bias=False)
    ql = pkg.QuantizedLinear.from_linear(lin)
    assert not ql.packed_weights.is_cuda
    assert sha(ql.packed_weights.numpy()) == str(Q["llama_up_packed_sha"])
    assert sha(ql.scales.numpy()) == str(Q["llama_up_scales_sha"])
    assert sha(ql.zero_points.numpy()) == str(Q["llama_up_zp_sha"])
    x = torch.randn(4096)
    y = ql.cuda()(x.cuda()).cpu().numpy()
    assert np.allclose(y, Q["llama_up_y"], atol=1e-3)


@pytest.mark.parametrize("N,K", [(1, 2), (3, 6), (5, 34), (64, 4096), (7, 11008)])
def test_quantize_random_vs_oracle(pkg, oracle, N, K):
    rng = np.random.default_rng(N * K)
    w = (rng.standard_normal((N, K)) * rng.uniform(0.01, 10, size=(N, 1))).astype(np.float32)
    p0, s0, z0 = oracle.quantize_weights(w)
    p, s, z = pkg.quantize_weights(torch.from_numpy(w).cuda())
    assert np.array_equal(p.cpu().numpy(), p0)
    assert np.array_equal(s.cpu().numpy(), s0) and np.array_equal(z.cpu().numpy(), z0)
    # encode -> decode -> encode is idempotent (the decoded grid re-quantises to the same codes)
    d = pkg.dequantize_weights(p, s, z)
    assert np.array_equal(d.cpu().numpy(), oracle.dequantize_weights(p0, s0, z0))


def test_quantize_asserts(pkg):
    with pytest.raises(AssertionError):
        pkg.quantize_weights(torch.randn(4, 7).cuda())
    with pytest.raises(AssertionError):
        pkg.quantize_weights(torch.randn(8).cuda())


def test_quantize_weights_moe_bit_exact(pkg):
    """python/moe_int4_module.py:19-80 (per-expert scalar scale / zero point)."""
    ws = [torch.from_numpy(w).cuda() for w in MI["w_fp16"]]
    p, s, z = pkg.quantize_weights_moe(ws)
    assert np.array_equal(p.cpu().numpy(), MI["packed"])
    assert np.array_equal(s.cpu().numpy(), MI["scales"])
    assert np.array_equal(z.cpu().numpy(), MI["zp"])


def test_quantize_rows_with_nan_and_inf_match_the_reference(pkg, oracle):
    """torch.min / max propagate NaN and the formulas of python/quantize.py:80-109 then run on non-finite numbers: a row
    with a NaN gets scale = zp = NaN, a row with +inf scale = inf and zp = 0, a row with -inf scale = inf and zp = NaN;
    every byte of such rows is 0.  (Checked against the real reference when this test was written; the oracle restates it.)"""
    import warnings
    rng = np.random.default_rng(5)
    w = rng.standard_normal((6, 64)).astype(np.float32)
    w[1, 3] = np.nan
    w[2, 5] = np.inf
    w[3, 7] = -np.inf
    w[4, :] = np.inf
    w[5, 0], w[5, 9] = np.inf, -np.inf
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        p0, s0, z0 = oracle.quantize_weights(w)
    p, s, z = (t.cpu().numpy() for t in pkg.quantize_weights(torch.from_numpy(w).cuda()))
    assert np.array_equal(p, p0) and not p[1:].any()
    assert np.array_equal(s, s0, equal_nan=True) and np.array_equal(z, z0, equal_nan=True)
    assert np.isnan(s[1]) and np.isnan(z[1]) and np.isinf(s[2]) and z[2] == 0 and np.isnan(z[3])
