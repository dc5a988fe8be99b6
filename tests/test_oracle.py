"""The numpy oracle (oracle/int4_oracle.py) against the golden outputs of the REAL reference
(tests/golden/*.npz, made by tests/golden/make_golden.py).  Runs on CPU."""
import hashlib
import os

import numpy as np
import pytest

from conftest import load_golden


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


Q = load_golden("quantize")
L = load_golden("linear")
R = load_golden("routing")
M = load_golden("moe")
MI = load_golden("moe_int4_module")


@pytest.mark.parametrize("name", ["s42_16x32", "s123_256x512", "s42_64x128", "s42_256x512", "edge"])
def test_quantize_bit_exact(oracle, name):
    p, s, z = oracle.quantize_weights(Q[f"{name}_w"])
    assert np.array_equal(p, Q[f"{name}_packed"])
    assert np.array_equal(s.view(np.uint32), Q[f"{name}_scales"].view(np.uint32))
    # zero points: values equal (torch yields -0.0 for an all-zero row; compare as numbers and signs)
    assert np.array_equal(z, Q[f"{name}_zp"])
    assert np.array_equal(np.signbit(z), np.signbit(Q[f"{name}_zp"]))


@pytest.mark.parametrize("name", ["s42_16x32", "s123_256x512", "s42_64x128", "s42_256x512", "edge"])
def test_dequantize_bit_exact(oracle, name):
    d = oracle.dequantize_weights(Q[f"{name}_packed"], Q[f"{name}_scales"], Q[f"{name}_zp"])
    assert np.array_equal(d.view(np.uint32), Q[f"{name}_deq"].view(np.uint32))


def test_quantize_4096_seed7_digest(oracle):
    """tests/test_correctness.py:68 / :236 inputs (regenerated from the torch seed)."""
    import torch
    torch.manual_seed(7)
    w = torch.randn(4096, 4096)
    x = torch.randn(4096)
    p, s, z = oracle.quantize_weights(w.numpy())
    assert sha(p) == str(Q["s7_4096_packed_sha"])
    assert int(p.astype(np.int64).sum()) == int(Q["s7_4096_packed_sum"])
    assert np.array_equal(s, Q["s7_4096_scales"]) and np.array_equal(z, Q["s7_4096_zp"])
    y = oracle.reference_quantized_linear(x.numpy(), p, s, z)
    # fp32 matmul: summation order differs between BLAS builds -> tolerance, same as the reference's 1e-2
    assert np.allclose(y, Q["s7_4096_y"], atol=1e-3, rtol=0)


def test_llama_up_digest(oracle):
    """BASELINE config 1: nn.Linear(4096, 11008, bias=False), seed 42 -> from_linear."""
    import torch
    torch.manual_seed(42)
    lin = torch.nn.Linear(4096, 11008, bias=False)
    p, s, z = oracle.quantize_weights(lin.weight.detach().numpy())
    assert sha(p) == str(Q["llama_up_packed_sha"])
    assert sha(s) == str(Q["llama_up_scales_sha"])
    assert sha(z) == str(Q["llama_up_zp_sha"])


def test_reference_linear_vectors(oracle):
    """The oracle outputs behind the reference's CUDA tests (tests/test_correctness.py:201-234)."""
    y = oracle.reference_quantized_linear(L["t1d_x"], Q["s42_64x128_packed"], Q["s42_64x128_scales"], Q["s42_64x128_zp"])
    assert np.allclose(y, L["t1d_y"], atol=1e-5, rtol=0)
    y = oracle.reference_quantized_linear(L["tb_x"], Q["s42_256x512_packed"], Q["s42_256x512_scales"], Q["s42_256x512_zp"])
    assert np.allclose(y, L["tb_y"], atol=1e-4, rtol=0)
    p, s, z = oracle.quantize_weights(L["odd_w"])
    assert np.array_equal(p, L["odd_packed"])
    y = oracle.reference_quantized_linear(L["odd_x"], p, s, z)
    assert np.allclose(y, L["odd_y"], atol=1e-5, rtol=0)


@pytest.mark.parametrize("dist", ["random", "skewed"])
def test_routing_indices_bit_exact(oracle, dist):
    idx, w = oracle.softmax_topk(R[f"{dist}64_logits"], 2)
    assert np.array_equal(idx, R[f"{dist}64_idx"])
    assert np.allclose(w, R[f"{dist}64_w"], atol=1e-6, rtol=0)
    counts, offsets = oracle.histogram_offsets(idx, 8)
    assert np.array_equal(counts, R[f"{dist}64_counts"])
    assert np.array_equal(offsets, R[f"{dist}64_offsets"])


@pytest.mark.parametrize("dist", ["random", "skewed"])
@pytest.mark.parametrize("T", [8192, 16384])
def test_routing_full_size(oracle, dist, T):
    """BASELINE configs 4/5: T tokens, E=8, k=2, seed 42, logits generated as routing.py:51-69 does."""
    import torch
    torch.manual_seed(42)
    if dist == "random":
        logits = torch.randn(T, 8)
    else:
        ep = 1.0 / (torch.arange(8, dtype=torch.float32) + 1)
        ep = ep / ep.sum()
        logits = torch.log(ep + 1e-10).unsqueeze(0).expand(T, -1)
        logits = logits + torch.randn_like(logits) * 0.5
    idx, w = oracle.softmax_topk(logits.numpy(), 2)
    assert sha(idx) == str(R[f"{dist}{T}_idx_sha"])
    counts, offsets = oracle.histogram_offsets(idx, 8)
    assert np.array_equal(counts, R[f"{dist}{T}_counts"])
    assert np.array_equal(offsets, R[f"{dist}{T}_offsets"])
    assert abs(float(w.astype(np.float64).sum()) - float(R[f"{dist}{T}_w_sum"])) < 1e-2


@pytest.mark.parametrize("dist", ["random", "skewed"])
def test_dispatch_combine(oracle, dist):
    idx = R[f"{dist}64_idx"]
    x = R[f"{dist}64_x"]
    xs, inverse = oracle.create_expert_inputs(x, idx, 8)
    for e in range(8):
        toks = sorted(int(np.where((x == row).all(axis=1))[0][0]) for row in xs[e])
        assert toks == list(R[f"{dist}64_expert{e}_tokens"])
    out = oracle.combine_expert_outputs(xs, R[f"{dist}64_w"], inverse, 2)
    assert np.allclose(out, R[f"{dist}64_combined_identity"], atol=1e-6, rtol=0)


def test_moe_layers(oracle):
    E = M["packed"].shape[0]
    for e in range(E):
        p, s, z = oracle.quantize_weights(M["w_fp16"][e].astype(np.float32))
        assert np.array_equal(p, M["packed"][e])
        assert np.array_equal(s, M["scales"][e]) and np.array_equal(z, M["zp"][e])
    experts = [(M["packed"][e], M["scales"][e], M["zp"][e]) for e in range(E)]
    idx, w = oracle.softmax_topk(M["logits"], 2)
    assert np.array_equal(idx, M["idx"])
    y = oracle.moe_single_projection(M["x"], M["logits"], experts, 2)
    assert np.allclose(y, M["y_single"], atol=2e-5, rtol=0)
    q = lambda ws: [oracle.quantize_weights(wi.astype(np.float32)) for wi in ws]
    y = oracle.moe_gated(M["x"], M["logits"], q(M["g_w1"]), q(M["g_w3"]), q(M["g_w2"]), 2)
    assert np.allclose(y, M["y_gated"], atol=2e-5, rtol=0)


def test_quantize_weights_moe(oracle):
    p, s, z = oracle.quantize_weights_moe([w for w in MI["w_fp16"]])
    assert np.array_equal(p, MI["packed"])
    assert np.array_equal(s, MI["scales"])
    assert np.array_equal(z, MI["zp"])


# ---- oracle/oracle.c (the timed CPU arm of bench.py when oracle/_ref is absent): pinned on the same golden vectors --
@pytest.fixture(scope="module")
def c_oracle():
    import subprocess
    sys_path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle")
    if not os.path.exists(os.path.join(sys_path, "_build", "liboracle.so")):
        subprocess.check_call(["make", "-C", sys_path, "-s"])
    import sys
    sys.path.insert(0, sys_path)
    import c_oracle as c
    return c


@pytest.mark.parametrize("name", ["s42_16x32", "s123_256x512", "s42_64x128", "s42_256x512", "edge"])
def test_c_oracle_quantize_dequantize_bit_exact(c_oracle, name):
    p, s, z = c_oracle.quantize_weights(Q[f"{name}_w"])
    assert np.array_equal(p, Q[f"{name}_packed"])
    assert np.array_equal(s.view(np.uint32), Q[f"{name}_scales"].view(np.uint32))
    assert np.array_equal(z, Q[f"{name}_zp"])
    d = c_oracle.dequantize_weights(Q[f"{name}_packed"], Q[f"{name}_scales"], Q[f"{name}_zp"])
    assert np.array_equal(d.view(np.uint32), Q[f"{name}_deq"].view(np.uint32))


def test_c_oracle_linear_vectors_and_numpy_oracle(c_oracle, oracle):
    """The C restatement against the reference's own linear vectors and, on a Llama-sized slice, against the numpy
    oracle (fp32 accumulation in a different order: tolerance, as in the reference's tests)."""
    y = c_oracle.Linear(Q["s42_64x128_packed"], Q["s42_64x128_scales"], Q["s42_64x128_zp"])(L["t1d_x"])
    assert np.allclose(y[0], L["t1d_y"], atol=1e-5, rtol=0)
    y = c_oracle.Linear(Q["s42_256x512_packed"], Q["s42_256x512_scales"], Q["s42_256x512_zp"])(L["tb_x"])
    assert np.allclose(y, L["tb_y"], atol=1e-4, rtol=0)
    rng = np.random.default_rng(0)
    packed = rng.integers(0, 256, size=(512, 2048), dtype=np.uint8)
    scales = (rng.random(512, dtype=np.float32) * 0.01 + 0.001).astype(np.float32)
    zps = rng.integers(0, 16, size=512).astype(np.float32)
    x = rng.standard_normal((3, 4096), dtype=np.float32)
    ref = oracle.reference_quantized_linear(x, packed, scales, zps, acc=np.float64)
    y = c_oracle.Linear(packed, scales, zps)(x)
    assert np.abs(y - ref).max() <= 2e-5 * np.abs(ref).max()
    assert c_oracle.max_threads() >= 1
