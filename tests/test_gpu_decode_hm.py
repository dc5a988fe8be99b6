"""Mid-batch decode kernel (csrc/gemv_hm.cu, force_path 7; the default for 3 <= M <= 16 when the rows of a CTA fit in
shared memory) against the float64 oracle, through the C ABI.

Tolerance: fp32 activations are carried as a 22-bit fixed-point number per (warp, 256 columns, token) in three IMMA
digits (default, hm_i3 = 1) or as fp16 hi + lo parts on HMMA (hm_i3 = 0), partial sums in s32 / fp32; measured max
error 0.8e-6 / 1.0e-6 of the largest output of a batch row (tools/hm_check.py); the bar here is 4e-6 per batch row.
16-bit activations (HMMA form) are exact in the operand (error = fp32 accumulation only)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def cuda(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _weights(rng, N, K):
    packed = rng.integers(0, 256, size=(N, K // 2), dtype=np.uint8)
    scales = (rng.random(N, dtype=np.float32) * 0.01 + 0.001).astype(np.float32)
    zps = rng.integers(0, 16, size=N).astype(np.float32)
    return packed, scales, zps


def _forced(pkg, path, fn, i3=-1):
    pkg._lib.tune("force_path", path)
    pkg._lib.tune("hm_i3", i3)
    try:
        return fn()
    finally:
        pkg._lib.tune("force_path", -1)
        pkg._lib.tune("hm_i3", -1)


@pytest.mark.parametrize("M", [1, 2, 3, 4, 5, 8, 9, 13, 16, 17, 24, 32])
@pytest.mark.parametrize("N,K", [(1, 256), (7, 256), (9, 512), (16, 768), (200, 1024), (2371, 2048), (11008, 4096), (4096, 4096),
                                 (3000, 6144), (4096, 11008), (1500, 8192), (600, 16384), (14336, 4096), (4096, 14336), (22016, 4096)])
def test_hm_kernel_edge_shapes(oracle, pkg, M, N, K):
    """Ragged row counts (fewer tiles than SMs, last tile partly foreign / out of bounds), one and two passes of eight
    tokens with a ragged last pass, one to four column pairs per warp, K < 4096 (warps without a pair), batch rows of
    one to four passes of eight tokens, very different magnitude (every (warp, pair, token) has its own scale), shapes that
    need two or more waves of CTAs
    (Mixtral's 14336-wide projections, Llama's gate + up rows in one matrix)."""
    rng = np.random.default_rng(1000 * M + N + K)
    packed, scales, zps = _weights(rng, N, K)
    x = (rng.standard_normal((M, K)) * rng.choice([1e-3, 1.0, 300.0], size=(M, 1))).astype(np.float32)
    P, S, Z, X = cuda(packed), cuda(scales), cuda(zps), cuda(x)
    rows = np.arange(N) if N <= 512 else rng.choice(N, size=512, replace=False)
    ref = oracle.reference_quantized_linear(x, packed[rows], scales[rows], zps[rows], acc=np.float64)
    y = _forced(pkg, 7, lambda: pkg._lib.linear_fwd(X, P, S, Z)).cpu().numpy()
    # the fp16 hi / lo HMMA form of the same kernel (what 16-bit activations always use) agrees to the same bar
    yh = _forced(pkg, 7, lambda: pkg._lib.linear_fwd(X, P, S, Z), i3=0).cpu().numpy()
    # (few outputs: max |y| of a row is not a scale -- a single output may be small by cancellation; sum_k |w x| is)
    absdot = np.abs(x).astype(np.float64) @ np.abs(oracle.dequantize_weights(packed[rows], scales[rows], zps[rows]).astype(np.float64)).T
    for m in range(M):
        for out in (y, yh):
            err = np.abs(ref[m] - out[m, rows])
            assert (err <= 4e-6 * np.abs(ref[m]).max() + 2e-7 * absdot[m]).all(), f"row {m}: {err.max()} vs |y|max {np.abs(ref[m]).max()}"
    # deterministic (fixed fold order), and exactly linear under power-of-two scaling of x
    y2 = _forced(pkg, 7, lambda: pkg._lib.linear_fwd(X, P, S, Z)).cpu().numpy()
    assert np.array_equal(y, y2)
    y4 = _forced(pkg, 7, lambda: pkg._lib.linear_fwd(4 * X, P, S, Z)).cpu().numpy()
    assert np.array_equal(y4, 4 * y)
    assert not _forced(pkg, 7, lambda: pkg._lib.linear_fwd(torch.zeros_like(X), P, S, Z)).cpu().numpy().any()


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("M,N,K", [(3, 300, 1024), (8, 11008, 4096), (16, 4096, 11008), (11, 1000, 2048), (27, 11008, 4096)])
def test_hm_kernel_16bit_activations(oracle, pkg, dtype, M, N, K):
    rng = np.random.default_rng(N + K + M)
    packed, scales, zps = _weights(rng, N, K)
    X = cuda(rng.standard_normal((M, K), dtype=np.float32)).to(dtype)
    P, S, Z = cuda(packed), cuda(scales), cuda(zps)
    rows = np.arange(N) if N <= 512 else rng.choice(N, size=512, replace=False)
    ref = oracle.reference_quantized_linear(X.float().cpu().numpy(), packed[rows], scales[rows], zps[rows], acc=np.float64)
    y = _forced(pkg, 7, lambda: pkg._lib.linear_fwd(X, P, S, Z, out_dtype=torch.float32)).cpu().numpy()
    assert np.abs(y[:, rows] - ref).max() <= 4e-6 * np.abs(ref).max()
    # 16-bit outputs: one more rounding
    yh = _forced(pkg, 7, lambda: pkg._lib.linear_fwd(X, P, S, Z)).float().cpu().numpy()
    eps = 2.0 ** -8 if dtype == torch.bfloat16 else 2.0 ** -11
    assert np.abs(yh[:, rows] - ref).max() <= eps * np.abs(ref).max()


@pytest.mark.parametrize("M", [3, 8, 12, 16, 21])
@pytest.mark.parametrize("F,K", [(150, 1024), (5504, 4096), (11008, 4096)])
def test_hm_kernel_gated_and_bias(oracle, pkg, M, F, K):
    """Fused gate + up pair (rows 2f / 2f+1 interleaved, h = silu(gate) * up) and the bias epilogue."""
    rng = np.random.default_rng(M + F + K)
    packed, scales, zps = _weights(rng, 2 * F, K)
    scales *= 0.3
    x = rng.standard_normal((M, K), dtype=np.float32)
    P, S, Z, X = cuda(packed), cuda(scales), cuda(zps), cuda(x)
    full = oracle.reference_quantized_linear(x, packed, scales, zps, acc=np.float64)
    ref = oracle.silu(full[:, 0::2]) * full[:, 1::2]
    h = _forced(pkg, 7, lambda: pkg._lib.linear_gated_fwd(X, P, S, Z)).cpu().numpy()
    assert h.shape == (M, F)
    assert np.abs(h - ref).max() <= 1e-5 * np.abs(ref).max()
    bias = rng.standard_normal(2 * F).astype(np.float32)
    y = _forced(pkg, 7, lambda: pkg._lib.linear_fwd(X, P, S, Z, bias=cuda(bias))).cpu().numpy()
    assert np.abs(y - (full + bias)).max() <= 4e-6 * np.abs(full).max()


@pytest.mark.parametrize("M", [3, 4, 9, 16, 29])
@pytest.mark.parametrize("dtype,i3", [(torch.float32, 1), (torch.float32, 0), (torch.bfloat16, 1)])
def test_hm_kernel_nonfinite_rows(oracle, pkg, M, dtype, i3):
    """python/quantize.py:172, 202: NaN / Inf in a row of x propagate as in dequantize + F.linear; the other rows of
    the batch are not disturbed."""
    rng = np.random.default_rng(M)
    N, K = 300, 1024
    packed, scales, zps = _weights(rng, N, K)
    x = rng.standard_normal((M, K), dtype=np.float32)
    x[0, 5] = np.nan
    x[1, 17] = np.inf
    x[1, 700] = -np.inf
    if M > 3:
        x[M - 1, 3] = -np.inf
    X = cuda(x).to(dtype)
    xr = X.float().cpu().numpy()
    with np.errstate(all="ignore"):
        ref = oracle.reference_quantized_linear(xr, packed, scales, zps)
    y = _forced(pkg, 7, lambda: pkg._lib.linear_fwd(X, cuda(packed), cuda(scales), cuda(zps), out_dtype=torch.float32), i3=i3).cpu().numpy()
    assert np.array_equal(np.isnan(y), np.isnan(ref))
    assert np.array_equal(np.isposinf(y), np.isposinf(ref)) and np.array_equal(np.isneginf(y), np.isneginf(ref))
    clean = [m for m in range(M) if np.isfinite(xr[m]).all()]
    ref64 = oracle.reference_quantized_linear(np.where(np.isfinite(xr), xr, 0.0), packed, scales, zps, acc=np.float64)
    assert np.abs(y[clean] - ref64[clean]).max() <= 4e-6 * np.abs(ref64[clean]).max()


def test_hm_kernel_is_the_default_for_mid_batches_and_agrees_with_the_integer_kernel(oracle, pkg):
    """Dispatch: 3 <= M <= 16 on a Llama shape takes the HMMA kernel (bit-identical to force_path 7); the exact-integer
    kernel (force_path 6) gives the same values to fp32 accumulation error."""
    rng = np.random.default_rng(5)
    N, K, M = 11008, 4096, 8
    packed, scales, zps = _weights(rng, N, K)
    X, P, S, Z = cuda(rng.standard_normal((M, K), dtype=np.float32)), cuda(packed), cuda(scales), cuda(zps)
    y = pkg._lib.linear_fwd(X, P, S, Z).cpu().numpy()
    y7 = _forced(pkg, 7, lambda: pkg._lib.linear_fwd(X, P, S, Z)).cpu().numpy()
    y6 = _forced(pkg, 6, lambda: pkg._lib.linear_fwd(X, P, S, Z)).cpu().numpy()
    assert np.array_equal(y, y7)
    assert np.abs(y7 - y6).max() <= 2e-6 * np.abs(y6).max()
    with pytest.raises(RuntimeError):                        # K must be a multiple of 256 (whole TMA boxes)
        _forced(pkg, 7, lambda: pkg._lib.linear_fwd(X[:, :384].contiguous(), cuda(np.zeros((64, 192), np.uint8)), cuda(np.ones(64, np.float32)), cuda(np.zeros(64, np.float32))))
