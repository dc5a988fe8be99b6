"""GPU parity tests for the fused dequantize-linear path, through the reference-facing API
(fused_quant_linear_cuda.forward / QuantizedLinear) which calls the C ABI of libb200q.so.

Tolerances: the reference's own (tests/test_correctness.py:201-253): atol 1e-3 for K <= 512,
atol 1e-2 for K = 4096 against the fp32 dequantize+matmul oracle; tighter bounds vs the float64
oracle are asserted too and documented per test."""
import numpy as np
import pytest
import torch

from conftest import load_golden

pytestmark = pytest.mark.gpu

Q = load_golden("quantize")
L = load_golden("linear")


def cuda(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def ext_forward(x, p, s, z):
    import fused_quant_linear_cuda
    return fused_quant_linear_cuda.forward(cuda(x), cuda(p), cuda(s), cuda(z)).cpu().numpy()


# ---- the reference's three CUDA tests, same seeds / shapes / tolerances -----------------------
def test_cuda_matches_reference_1d(oracle):
    torch.manual_seed(42)
    weight = torch.randn(64, 128)
    x = torch.randn(128)
    p, s, z = oracle.quantize_weights(weight.numpy())
    ref = oracle.reference_quantized_linear(x.numpy(), p, s, z)
    out = ext_forward(x.numpy(), p, s, z)
    assert out.shape == (64,)
    assert np.allclose(ref, out, atol=1e-3), f"Max diff: {np.abs(ref - out).max()}"
    assert np.allclose(L["t1d_y"], out, atol=1e-3)          # the real reference's output


def test_cuda_matches_reference_batched(oracle):
    torch.manual_seed(42)
    weight = torch.randn(256, 512)
    x = torch.randn(4, 512)
    p, s, z = oracle.quantize_weights(weight.numpy())
    ref = oracle.reference_quantized_linear(x.numpy(), p, s, z)
    out = ext_forward(x.numpy(), p, s, z)
    assert out.shape == (4, 256)
    assert np.allclose(ref, out, atol=1e-3), f"Max diff: {np.abs(ref - out).max()}"
    assert np.allclose(L["tb_y"], out, atol=1e-3)


def test_cuda_large_dims(oracle):
    torch.manual_seed(7)
    weight = torch.randn(4096, 4096)
    x = torch.randn(4096)
    p, s, z = oracle.quantize_weights(weight.numpy())
    ref = oracle.reference_quantized_linear(x.numpy(), p, s, z)
    out = ext_forward(x.numpy(), p, s, z)
    assert np.allclose(ref, out, atol=1e-2), f"Max diff: {np.abs(ref - out).max()}"
    assert np.allclose(Q["s7_4096_y"], out, atol=1e-2)
    ref64 = oracle.reference_quantized_linear(x.numpy(), p, s, z, acc=np.float64)
    err = np.abs(ref64 - out).max()
    print(f"4096x4096 max abs err vs float64 oracle: {err:.3e} (|y|max {np.abs(ref64).max():.1f})")
    assert err < 2e-3


# ---- shape / batch sweep over every kernel path ------------------------------------------------
@pytest.mark.parametrize("N,K", [(64, 128), (100, 256), (256, 512), (1000, 1024), (37, 90), (8, 2), (129, 384)])
@pytest.mark.parametrize("M", [1, 2, 3, 4, 5, 8, 13, 16, 17, 40])
def test_linear_sweep(oracle, N, K, M):
    rng = np.random.default_rng(N * 131 + K * 7 + M)
    w = rng.standard_normal((N, K), dtype=np.float32)
    x = rng.standard_normal((M, K), dtype=np.float32)
    p, s, z = oracle.quantize_weights(w)
    ref = oracle.reference_quantized_linear(x, p, s, z, acc=np.float64)
    out = ext_forward(x, p, s, z)
    assert out.shape == (M, N)
    tol = 1e-3 if K <= 512 else 2e-3
    assert np.abs(ref - out).max() < tol, f"max err {np.abs(ref - out).max()}"


@pytest.mark.parametrize("K,N", [(4096, 11008), (11008, 4096), (4096, 14336), (14336, 4096)])
@pytest.mark.parametrize("M", [1, 4, 16])
def test_llama_mixtral_decode_shapes(oracle, K, N, M):
    """BASELINE config 2 shapes at full size: rows sampled against the float64 oracle plus two
    size-independent properties: linearity in x and exact zero for zero input."""
    rng = np.random.default_rng(K + N + M)
    packed = rng.integers(0, 256, size=(N, K // 2), dtype=np.uint8)
    scales = (rng.random(N, dtype=np.float32) * 0.01 + 0.001).astype(np.float32)
    zps = rng.integers(0, 16, size=N).astype(np.float32)
    x = rng.standard_normal((M, K), dtype=np.float32)
    import fused_quant_linear_cuda as ext
    P, S, Z = cuda(packed), cuda(scales), cuda(zps)
    y = ext.forward(cuda(x), P, S, Z).cpu().numpy()
    rows = rng.choice(N, size=256, replace=False)
    ref = oracle.reference_quantized_linear(x, packed[rows], scales[rows], zps[rows], acc=np.float64)
    err = np.abs(ref - y[:, rows]).max()
    scale = np.abs(ref).max()
    print(f"K={K} N={N} M={M}: max abs err {err:.3e}, |y|max {scale:.2f}, rel {err / scale:.2e}")
    # exact-integer IMMA path (error = final fp32 rounding) for M <= 16 on the Llama shapes and M <= 8 on Mixtral's
    # (resident decode kernel, ring of tile buffers); Mixtral M = 16: tcgen05 path, fp16 hi/lo split of x, fp32 accumulation over up to 14336 terms
    assert err < 1e-2 and err / scale < (2e-5 if (M <= 8 or 14336 not in (K, N)) else 1e-4)
    # linearity: f(2x) == 2 f(x) exactly (power-of-two scaling commutes with every rounding step)
    y2 = ext.forward(cuda(2 * x), P, S, Z).cpu().numpy()
    assert np.array_equal(y2, 2 * y)
    # determinism: bit-identical on a second run
    assert np.array_equal(ext.forward(cuda(x), P, S, Z).cpu().numpy(), y)
    assert not np.any(ext.forward(cuda(np.zeros_like(x)), P, S, Z).cpu().numpy())


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
@pytest.mark.parametrize("K,N,M", [(4096, 11008, 300), (11008, 4096, 513), (4096, 1000, 64)])
def test_prefill_tcgen05_shapes(oracle, pkg, dtype, K, N, M):
    """BASELINE config 3 shapes (tcgen05 path, M > 8): sampled weight rows against the float64 oracle,
    ragged M (not a multiple of the 256-token tile) and N (not a multiple of the 128-row tile).
    Tolerances: fp32 activations keep the reference's atol 1e-2 (hi/lo split, ~1e-6 relative);
    16-bit activations are exact in the products, so only the output rounding of that type remains."""
    rng = np.random.default_rng(K + N + M)
    packed = rng.integers(0, 256, size=(N, K // 2), dtype=np.uint8)
    scales = (rng.random(N, dtype=np.float32) * 0.01 + 0.001).astype(np.float32)
    zps = rng.integers(0, 16, size=N).astype(np.float32)
    x = torch.from_numpy(rng.standard_normal((M, K), dtype=np.float32)).to(dtype)
    y = pkg._lib.linear_fwd(x.cuda(), cuda(packed), cuda(scales), cuda(zps))
    assert y.dtype == dtype and y.shape == (M, N)
    rows = rng.choice(N, size=96, replace=False)
    ref = oracle.reference_quantized_linear(x.float().numpy(), packed[rows], scales[rows], zps[rows], acc=np.float64)
    got = y.float().cpu().numpy()[:, rows]
    err = np.abs(ref - got).max()
    scale = np.abs(ref).max()
    print(f"K={K} N={N} M={M} {dtype}: max abs err {err:.3e}, |y|max {scale:.2f}, rel {err / scale:.2e}")
    eps = {torch.float32: 1e-4, torch.float16: 2.0 ** -10, torch.bfloat16: 2.0 ** -7}[dtype]
    assert err <= eps * scale + 1e-4
    if dtype == torch.float32:
        assert err < 1e-2
    # row independence: the first 5 token rows alone give the same values (other tile shape / path)
    y5 = pkg._lib.linear_fwd(x[:5].cuda().contiguous(), cuda(packed), cuda(scales), cuda(zps)).float().cpu().numpy()
    assert np.abs(y5[:, rows] - ref[:5]).max() <= eps * scale + 1e-4


def test_forced_generic_path_matches(oracle, pkg):
    """The SIMT path mirrors the reference's arithmetic order (w = (q - zp) * s first)."""
    rng = np.random.default_rng(0)
    w = rng.standard_normal((96, 256), dtype=np.float32)
    x = rng.standard_normal((3, 256), dtype=np.float32)
    p, s, z = oracle.quantize_weights(w)
    ref = oracle.reference_quantized_linear(x, p, s, z, acc=np.float64)
    pkg._lib.tune("force_path", 1)
    try:
        out = ext_forward(x, p, s, z)
    finally:
        pkg._lib.tune("force_path", -1)
    assert np.abs(ref - out).max() < 1e-4


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("M", [1, 7, 33])
def test_half_precision_activations(oracle, pkg, dtype, M):
    """Superset of the reference API: fp16 / bf16 activations and outputs (what QuantizedMoEExpert
    feeds, moe_int4_module.py:71-72).  Error bound: output rounding of the 16-bit type."""
    rng = np.random.default_rng(M)
    w = (rng.standard_normal((192, 512)) * 0.05).astype(np.float32)
    p, s, z = oracle.quantize_weights(w)
    x = torch.from_numpy(rng.standard_normal((M, 512)).astype(np.float32)).to(dtype)
    ref = oracle.reference_quantized_linear(x.float().numpy(), p, s, z, acc=np.float64)
    y = pkg._lib.linear_fwd(x.cuda(), cuda(p), cuda(s), cuda(z))
    assert y.dtype == dtype
    eps = 2.0 ** -10 if dtype == torch.float16 else 2.0 ** -7
    assert np.abs(ref - y.float().cpu().numpy()).max() <= eps * np.abs(ref).max() + 1e-4


def test_module_forward_and_state_dict(oracle, pkg):
    torch.manual_seed(42)
    lin = torch.nn.Linear(512, 256, bias=False)
    ql = pkg.QuantizedLinear.from_linear(lin.cuda())
    p, s, z = oracle.quantize_weights(lin.weight.detach().cpu().numpy())
    assert np.array_equal(ql.packed_weights.cpu().numpy(), p)
    x = torch.randn(2, 3, 512, device="cuda")
    y = ql(x)
    assert y.shape == (2, 3, 256)
    ref = oracle.reference_quantized_linear(x.reshape(-1, 512).cpu().numpy(), p, s, z, acc=np.float64)
    assert np.abs(ref - y.reshape(-1, 256).cpu().numpy()).max() < 1e-3
    y1 = ql(x[0, 0])                      # second call takes the static-weights (PDL) launch
    assert y1.shape == (256,)
    assert np.abs(ref[0] - y1.cpu().numpy()).max() < 1e-3
    ql2 = pkg.QuantizedLinear(512, 256).cuda()
    ql2.load_state_dict(ql.state_dict())
    assert torch.equal(ql2(x), y)
    with pytest.raises(RuntimeError):
        ql(x.cpu())


def test_extension_error_behaviour(pkg):
    import fused_quant_linear_cuda as ext
    p = torch.zeros(8, 16, dtype=torch.uint8, device="cuda")
    s = torch.ones(8, device="cuda")
    z = torch.zeros(8, device="cuda")
    x = torch.zeros(32, device="cuda")
    assert ext.forward(x, p, s, z).shape == (8,)
    with pytest.raises(RuntimeError, match="input must be float32"):
        ext.forward(x.half(), p, s, z)
    with pytest.raises(RuntimeError, match="packed_weights must be uint8"):
        ext.forward(x, p.float(), s, z)
    with pytest.raises(RuntimeError, match="packed_weights dim 1 must be input_dim / 2"):
        ext.forward(torch.zeros(30, device="cuda"), p, s, z)
    with pytest.raises(RuntimeError, match="input must be contiguous"):
        ext.forward(torch.zeros(4, 64, device="cuda")[:, ::2], p, s, z)
    with pytest.raises(RuntimeError, match="packed_weights must be a CUDA tensor"):
        ext.forward(x, p.cpu(), s, z)
    assert ext.forward(torch.zeros(0, 32, device="cuda"), p, s, z).shape == (0, 8)


def test_forward_host_pinned(oracle, pkg):
    """Host-resident caller: pinned x in, pinned y out through b200q_linear_fwd_host (one C-ABI call)."""
    torch.manual_seed(3)
    lin = torch.nn.Linear(1024, 768, bias=False)
    ql = pkg.QuantizedLinear.from_linear(lin.cuda())
    p, s, z = oracle.quantize_weights(lin.weight.detach().cpu().numpy())
    for M in (1, 5, 40):
        x = torch.randn(M, 1024).pin_memory()
        y = ql(x)                                     # CPU pinned tensor -> forward_host
        assert not y.is_cuda and y.shape == (M, 768)
        torch.cuda.synchronize()
        ref = oracle.reference_quantized_linear(x.numpy(), p, s, z, acc=np.float64)
        assert np.abs(ref - y.numpy()).max() < 1e-3
    with pytest.raises(RuntimeError):
        ql(torch.randn(2, 1024))                      # pageable host memory: no silent slow path


# ---- decode kernels: CTA-resident (gemv_dec.cu, force_path 6) and ring (gemv.cu, force_path 2, M <= 8) ----------
@pytest.mark.parametrize("M", [1, 2, 3, 4, 5, 8, 9, 13, 16])
@pytest.mark.parametrize("N,K", [(1, 128), (7, 128), (9, 256), (16, 384), (200, 1024), (2371, 2048), (11008, 4096),
                                 (3000, 6144), (4096, 11008), (1500, 8192), (600, 16384), (14336, 4096), (4096, 14336),
                                 (5000, 8192)])
def test_resident_decode_kernel_edge_shapes(oracle, pkg, M, N, K):
    """Ragged row counts (fewer tiles than SMs, last tile partly foreign / out of bounds), every batch size of the
    decode path (passes of two or four batch rows), one to four column pairs per warp, K = 128 mod 256 (half-empty
    last TMA box), shapes whose rows do not fit in a CTA's tile buffers (ring: buffers are refilled, outputs window by window); the
    resident kernel against the float64 oracle and, for M <= 8, against the ring kernel."""
    rng = np.random.default_rng(1000 * M + N + K)
    packed = rng.integers(0, 256, size=(N, K // 2), dtype=np.uint8)
    scales = (rng.random(N, dtype=np.float32) * 0.01 + 0.001).astype(np.float32)
    zps = rng.integers(0, 16, size=N).astype(np.float32)
    x = (rng.standard_normal((M, K)) * rng.choice([1e-3, 1.0, 300.0], size=(M, 1))).astype(np.float32)
    P, S, Z, X = cuda(packed), cuda(scales), cuda(zps), cuda(x)
    rows = np.arange(N) if N <= 512 else rng.choice(N, size=512, replace=False)
    ref = oracle.reference_quantized_linear(x, packed[rows], scales[rows], zps[rows], acc=np.float64)
    outs = {}
    for path in (6, 2) if M <= 8 else (6,):
        pkg._lib.tune("force_path", path)
        try:
            outs[path] = pkg._lib.linear_fwd(X, P, S, Z).cpu().numpy()
        except RuntimeError:
            # the ring kernel does not take every shape; the resident kernel wants whole 256-column TMA boxes
            assert path == 2 or K % 256 != 0
            continue
        finally:
            pkg._lib.tune("force_path", -1)
        for m in range(M):          # rows have very different magnitudes: the bar is per batch row
            err = np.abs(ref[m] - outs[path][m, rows]).max()
            assert err <= 1e-6 * np.abs(ref[m]).max() + 1e-30, f"path {path} row {m}: {err} vs |y|max {np.abs(ref[m]).max()}"
    if 2 in outs and 6 in outs:
        # the two kernels quantise the odd columns of x on different grids: equal to ~1e-7, not bit-equal
        for m in range(M):
            assert np.abs(outs[6][m] - outs[2][m]).max() <= 1e-6 * np.abs(outs[2][m]).max() + 1e-30
    # default dispatch (resident kernel when K % 256 == 0, else ring kernel / tcgen05): deterministic, within the bar
    y = pkg._lib.linear_fwd(X, P, S, Z).cpu().numpy()
    assert np.array_equal(pkg._lib.linear_fwd(X, P, S, Z).cpu().numpy(), y)
    # (wide K goes to the tcgen05 GEMM by default with M > 8, and to the ring kernel with M >= 3 when the CTA has fewer
    # tile buffers than tiles)
    # (3 <= M <= 16 with every tile resident: the HMMA kernel, tests/test_gpu_decode_hm.py -- equal to accumulation error)
    if 6 in outs and M <= 2:
        assert np.array_equal(y, outs[6])
    elif 6 in outs:
        for m in range(M):
            assert np.abs(y[m] - outs[6][m]).max() <= 1e-4 * np.abs(outs[6][m]).max() + 1e-30
    for m in range(M):
        assert np.abs(ref[m] - y[m, rows]).max() <= 1e-4 * np.abs(ref[m]).max() + 1e-30


@pytest.mark.parametrize("M", [1, 2, 3, 4, 7, 16])
@pytest.mark.parametrize("N,K", [(2371, 2048), (11008, 4096), (4096, 11008), (4736, 1024), (4737, 1024), (9000, 256), (600, 16384)])
@pytest.mark.parametrize("bufs", [1, 2, 3])
def test_resident_decode_kernel_ring_of_tile_buffers(oracle, pkg, M, N, K, bufs):
    """A CTA with fewer tile buffers than tiles (Mixtral's 14336-wide projections; forced here on smaller shapes with
    the tuning key gemv_bufs) refills a buffer as soon as every warp is done with it and produces its outputs window by
    window.  Integer partial sums: bit-identical to the all-resident form, whatever the number of buffers; plain and
    gated (silu(gate) * up) outputs."""
    rng = np.random.default_rng(M * 7 + N + K + bufs)
    packed = rng.integers(0, 256, size=(N - (N & 1), K // 2), dtype=np.uint8)
    N = packed.shape[0]
    scales = (rng.random(N, dtype=np.float32) * 0.01 + 0.001).astype(np.float32)
    zps = rng.integers(0, 16, size=N).astype(np.float32)
    x = rng.standard_normal((M, K), dtype=np.float32)
    P, S, Z, X = cuda(packed), cuda(scales), cuda(zps), cuda(x)
    pkg._lib.tune("force_path", 6)
    try:
        y0 = pkg._lib.linear_fwd(X, P, S, Z).cpu().numpy()
        g0 = pkg._lib.linear_gated_fwd(X, P, S, Z).cpu().numpy()
        pkg._lib.tune("gemv_bufs", bufs)
        y1 = pkg._lib.linear_fwd(X, P, S, Z).cpu().numpy()
        g1 = pkg._lib.linear_gated_fwd(X, P, S, Z).cpu().numpy()
    finally:
        pkg._lib.tune("gemv_bufs", -1)
        pkg._lib.tune("force_path", -1)
    assert np.array_equal(y0, y1) and np.array_equal(g0, g1)
    rows = rng.choice(N, size=min(N, 256), replace=False)
    ref = oracle.reference_quantized_linear(x, packed[rows], scales[rows], zps[rows], acc=np.float64)
    assert np.abs(ref - y1[:, rows]).max() <= 1e-6 * np.abs(ref).max()


@pytest.mark.parametrize("M,N,K", [(1, 11008, 4096), (2, 4096, 11008), (2, 700, 2048), (1, 2368, 4096), (1, 2369, 4096)])
def test_resident_decode_kernel_slot_and_pipelined_reductions_agree(oracle, pkg, M, N, K):
    """M <= 2 in one pass: per-(tile, warp) slots folded after the loop (default) against the pipelined double-buffer
    reduction the larger batches use (tuning key gemv_slots = 0).  Integer sums: bit-identical.  N = 2368 = 148 x 16:
    the row of 0x11 bytes that yields sum_k X needs a tile of its own."""
    rng = np.random.default_rng(M + N + K)
    packed = rng.integers(0, 256, size=(N, K // 2), dtype=np.uint8)
    scales = (rng.random(N, dtype=np.float32) * 0.01 + 0.001).astype(np.float32)
    zps = rng.integers(0, 16, size=N).astype(np.float32)
    x = rng.standard_normal((M, K), dtype=np.float32)
    ref = oracle.reference_quantized_linear(x, packed, scales, zps, acc=np.float64)
    outs = []
    for slots in (1, 0):
        pkg._lib.tune("gemv_slots", slots)
        try:
            outs.append(pkg._lib.linear_fwd(cuda(x), cuda(packed), cuda(scales), cuda(zps)).cpu().numpy())
        finally:
            pkg._lib.tune("gemv_slots", -1)
        assert np.abs(outs[-1] - ref).max() <= 1e-6 * np.abs(ref).max()
    assert np.array_equal(outs[0], outs[1])


def test_next_layer_hint_does_not_change_results(oracle, pkg):
    """b200q_linear_fwd_next / QuantizedLinear.set_next: an L2 prefetch hint only."""
    rng = np.random.default_rng(5)
    N, K = 11008, 4096
    layers = []
    for i in range(3):
        m = pkg.QuantizedLinear(K, N).cuda()
        m.packed_weights = cuda(rng.integers(0, 256, size=(N, K // 2), dtype=np.uint8))
        m.scales = cuda((rng.random(N, dtype=np.float32) * 0.01 + 0.001).astype(np.float32))
        m.zero_points = cuda(rng.integers(0, 16, size=N).astype(np.float32))
        layers.append(m)
    x = cuda(rng.standard_normal((1, K), dtype=np.float32))
    plain = [m(x).cpu().numpy() for m in layers]
    plain = [m(x).cpu().numpy() for m in layers]            # second pass: static-weights fast path
    pkg.link_decode_order(layers)                       # layers[i].set_next(layers[i + 1]), last -> first
    hinted = [m(x).cpu().numpy() for m in layers]
    for a, b in zip(plain, hinted):
        assert np.array_equal(a, b)
    # a hint that points at something unrelated (or misaligned) is harmless too
    y = pkg._lib.linear_fwd(x, layers[0].packed_weights, layers[0].scales, layers[0].zero_points,
                            next_packed=layers[1].packed_weights.view(-1)[1:4097])
    assert np.array_equal(y.cpu().numpy(), plain[0])


def test_decode_non_integer_zero_points(oracle, pkg):
    """Zero points given by the caller (b200q_quantize_rows_given) need not be integers: the epilogue then takes its
    fp64 branch."""
    rng = np.random.default_rng(11)
    N, K, M = 300, 1024, 2
    packed = rng.integers(0, 256, size=(N, K // 2), dtype=np.uint8)
    scales = (rng.random(N, dtype=np.float32) * 0.01 + 0.001).astype(np.float32)
    zps = (rng.random(N, dtype=np.float32) * 15).astype(np.float32)
    x = rng.standard_normal((M, K), dtype=np.float32)
    ref = oracle.reference_quantized_linear(x, packed, scales, zps, acc=np.float64)
    out = pkg._lib.linear_fwd(cuda(x), cuda(packed), cuda(scales), cuda(zps)).cpu().numpy()
    assert np.abs(ref - out).max() <= 2e-5 * np.abs(ref).max()


# ---- tcgen05 GEMM: stream-K and small token tiles ----------------------------------------------------------
@pytest.mark.parametrize("M,N,K", [(300, 4096, 11008), (70, 1000, 1024), (520, 2048, 2048), (1100, 640, 4096),
                                   (17, 11008, 4096), (33, 11008, 4096), (64, 14336, 4096), (24, 2000, 2048)])
def test_stream_k_gemm_matches_whole_tiles_and_is_deterministic(oracle, pkg, M, N, K):
    """Stream-K (forced with gemm_sk=1, also on shapes the heuristic would leave alone: several contributors per
    tile, tiles finished by a CTA that also publishes) against the float64 oracle and against the whole-tile schedule;
    a second run must be bit-identical (partials are added in CTA order) and must find the flags zeroed."""
    rng = np.random.default_rng(M + N + K)
    packed = rng.integers(0, 256, size=(N, K // 2), dtype=np.uint8)
    scales = (rng.random(N, dtype=np.float32) * 0.01 + 0.001).astype(np.float32)
    zps = rng.integers(0, 16, size=N).astype(np.float32)
    x = rng.standard_normal((M, K), dtype=np.float32)
    P, S, Z = cuda(packed), cuda(scales), cuda(zps)
    X = cuda(x).to(torch.bfloat16)
    ref = oracle.reference_quantized_linear(X.float().cpu().numpy(), packed, scales, zps, acc=np.float64)
    outs = {}
    for sk in (0, 1):
        pkg._lib.tune("force_path", 3)
        pkg._lib.tune("gemm_sk", sk)
        try:
            a = pkg._lib.linear_fwd(X, P, S, Z, out_dtype=torch.float32).cpu().numpy()
            b = pkg._lib.linear_fwd(X, P, S, Z, out_dtype=torch.float32).cpu().numpy()
        finally:
            pkg._lib.tune("gemm_sk", -1)
            pkg._lib.tune("force_path", -1)
        assert np.array_equal(a, b), f"gemm_sk={sk}: not deterministic"
        assert np.abs(a - ref).max() <= 1e-4 * np.abs(ref).max()
        outs[sk] = a
    # different summation split of the fp32 accumulation: equal to rounding, not bit-equal
    assert np.abs(outs[0] - outs[1]).max() <= 2e-5 * np.abs(ref).max()
    # the decode kernels share the first 4 KB of the workspace with the stream-K flags: they must still be zero
    y1 = pkg._lib.linear_fwd(X[:1].float().contiguous(), P, S, Z).cpu().numpy()
    assert np.abs(y1 - ref[:1]).max() <= 2e-5 * np.abs(ref[:1]).max()


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("M,N,K", [(4096, 128, 8192), (2048, 256, 11008), (1024, 384, 14336)])
def test_stream_k_few_long_tiles(oracle, pkg, dtype, M, N, K):
    """Few weight rows, many tokens, long K: the heuristic itself picks stream-K and every tile is cut between ~10 CTAs.
    Default dispatch and forced stream-K against the float64 oracle, fp32 and bf16 activations, three runs each (the
    fp32 instance with 256-token tiles -- three pipeline stages -- gave 0.3 relative error here: a dequant group skipped the
    other group's k-blocks without observing their fills, and its parity wait fell two phases behind)."""
    rng = np.random.default_rng(M + N + K)
    packed = rng.integers(0, 256, size=(N, K // 2), dtype=np.uint8)
    scales = (rng.random(N, dtype=np.float32) * 0.01 + 0.001).astype(np.float32)
    zps = rng.integers(0, 16, size=N).astype(np.float32)
    x = rng.standard_normal((M, K), dtype=np.float32)
    P, S, Z = cuda(packed), cuda(scales), cuda(zps)
    X = cuda(x).to(dtype)
    ref = oracle.reference_quantized_linear(X.float().cpu().numpy(), packed, scales, zps, acc=np.float64)
    for sk in (-1, 1):
        pkg._lib.tune("force_path", 3)
        pkg._lib.tune("gemm_sk", sk)
        try:
            for _ in range(3):
                y = pkg._lib.linear_fwd(X, P, S, Z, out_dtype=torch.float32).cpu().numpy()
                assert np.abs(y - ref).max() <= 1e-4 * np.abs(ref).max(), f"gemm_sk={sk}"
        finally:
            pkg._lib.tune("gemm_sk", -1)
            pkg._lib.tune("force_path", -1)


@pytest.mark.parametrize("bn", [32, 64, 128, 192, 256])
def test_gemm_token_tile_heights(oracle, pkg, bn):
    """Every token-tile height of the tcgen05 GEMM on one ragged shape (M, N not multiples of the tile)."""
    rng = np.random.default_rng(bn)
    M, N, K = 333, 777, 1536
    packed = rng.integers(0, 256, size=(N, K // 2), dtype=np.uint8)
    scales = (rng.random(N, dtype=np.float32) * 0.01 + 0.001).astype(np.float32)
    zps = rng.integers(0, 16, size=N).astype(np.float32)
    x = rng.standard_normal((M, K), dtype=np.float32)
    ref = oracle.reference_quantized_linear(x, packed, scales, zps, acc=np.float64)
    pkg._lib.tune("force_path", 3)
    pkg._lib.tune("gemm_bn", bn)
    try:
        y = pkg._lib.linear_fwd(cuda(x), cuda(packed), cuda(scales), cuda(zps)).cpu().numpy()
    finally:
        pkg._lib.tune("gemm_bn", -1)
        pkg._lib.tune("force_path", -1)
    assert np.abs(y - ref).max() <= 1e-4 * np.abs(ref).max()


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
@pytest.mark.parametrize("M,N,K", [(1, 4096, 11008), (2, 4100, 11008), (1, 37, 8192), (2, 1000, 7168), (7, 4096, 11008),
                                   (16, 4096, 11008), (12, 11008, 4096)])
def test_wide_k_and_16bit_decode(oracle, pkg, dtype, M, N, K):
    """K up to 11008 without a K split (three column pairs per warp; M > 2 in passes of two batch rows), ragged row
    counts, 16-bit activations, determinism; the ring kernel as a second opinion where it applies (M <= 8)."""
    rng = np.random.default_rng(N + K + M)
    packed = rng.integers(0, 256, size=(N, K // 2), dtype=np.uint8)
    scales = (rng.random(N, dtype=np.float32) * 0.01 + 0.001).astype(np.float32)
    zps = rng.integers(0, 16, size=N).astype(np.float32)
    X = cuda(rng.standard_normal((M, K), dtype=np.float32)).to(dtype)
    P, S, Z = cuda(packed), cuda(scales), cuda(zps)
    ref = oracle.reference_quantized_linear(X.float().cpu().numpy(), packed, scales, zps, acc=np.float64)
    outs = {}
    for path in (6, 2) if M <= 8 else (6,):
        pkg._lib.tune("force_path", path)
        try:
            outs[path] = pkg._lib.linear_fwd(X, P, S, Z, out_dtype=torch.float32).cpu().numpy()
        except RuntimeError:
            assert path == 2
            continue
        finally:
            pkg._lib.tune("force_path", -1)
        assert np.abs(outs[path] - ref).max() <= 1e-6 * np.abs(ref).max()
    if M <= 2:                                           # (3 <= M <= 16: the default is the HMMA kernel, tests/test_gpu_decode_hm.py)
        assert np.array_equal(pkg._lib.linear_fwd(X, P, S, Z, out_dtype=torch.float32).cpu().numpy(), outs[6])


def test_forward_host_with_bias_and_compiled_binding(oracle, pkg):
    """Pinned host activations through a module with a bias (device path + copies) and without one (one C call through
    the compiled binding when it is built): both against the float64 oracle."""
    torch.manual_seed(3)
    for bias in (False, True):
        lin = torch.nn.Linear(1024, 384, bias=bias)
        ql = pkg.QuantizedLinear.from_linear(lin.cuda())
        p0, s0, z0 = oracle.quantize_weights(lin.weight.detach().cpu().numpy())
        for M in (1, 5, 40):
            x = torch.randn(M, 1024).pin_memory()
            y = ql(x)
            torch.cuda.synchronize()
            assert not y.is_cuda and y.shape == (M, 384)
            ref = oracle.reference_quantized_linear(x.numpy(), p0, s0, z0, acc=np.float64)
            if bias:
                ref = ref + lin.bias.detach().cpu().numpy().astype(np.float64)
            assert np.abs(y.numpy() - ref).max() <= 1e-4 * np.abs(ref).max()


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_forward_host_direct_and_copy_paths_agree(oracle, pkg, dtype):
    """b200q_linear_fwd_host for decode-sized calls: staging kernel + stores straight into the pinned result buffer
    (default) against the two-cudaMemcpyAsync form (host_direct=0); also inside a CUDA graph, replayed with new x."""
    torch.manual_seed(5)
    lin = torch.nn.Linear(4096, 1000, bias=False)
    ql = pkg.QuantizedLinear.from_linear(lin.cuda())
    outs = {}
    for M in (1, 8):
        x = torch.randn(M, 4096).to(dtype).pin_memory()
        for hd in (1, 0):
            pkg._lib.tune("host_direct", hd)
            try:
                y = ql.forward_host(x)
                torch.cuda.synchronize()
                outs[hd] = y.float().numpy().copy()
            finally:
                pkg._lib.tune("host_direct", -1)
        assert np.array_equal(outs[0], outs[1])
        ref = ql(x.cuda()).float().cpu().numpy()
        assert np.array_equal(outs[1], ref)
    # graph capture of the direct path: replays must pick up the CURRENT contents of the pinned input
    x = torch.randn(1, 4096).to(dtype).pin_memory()
    out = torch.empty(1, 1000, dtype=dtype).pin_memory()
    ql.forward_host(x, out=out)
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        ql.forward_host(x, out=out)
    torch.cuda.current_stream().wait_stream(side)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        ql.forward_host(x, out=out)
    x.copy_(torch.randn(1, 4096).to(dtype))
    g.replay()
    torch.cuda.synchronize()
    assert np.array_equal(out.float().numpy(), ql(x.cuda()).float().cpu().numpy())
