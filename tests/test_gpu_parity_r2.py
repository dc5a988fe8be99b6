"""Round-2 parity cases: non-finite activations on every path, bias, the zero-fill contract of the grouped entry
point, router safety on NaN logits, stale stacked-weight caches.  All through the C ABI (ctypes)."""
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


def _same_nonfinite(got, ref):
    """NaN where the reference has NaN, +-Inf (same sign) where it has Inf, finite elsewhere."""
    assert np.array_equal(np.isnan(got), np.isnan(ref))
    assert np.array_equal(np.isposinf(got), np.isposinf(ref))
    assert np.array_equal(np.isneginf(got), np.isneginf(ref))


# M = 1, 2, 5, 16: resident decode kernel (one / two n-tiles, passes); 40: tcgen05 GEMM; force 2: ring; 1: generic
@pytest.mark.parametrize("M,force", [(1, -1), (2, -1), (5, -1), (16, -1), (40, -1), (300, -1), (3, 2), (4, 1)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_nonfinite_activations_match_dequantize_plus_linear(oracle, pkg, M, force, dtype):
    """python/quantize.py:172, 202: y = F.linear(x, (q - zp) * s).  A NaN in a row of x makes every output of that row
    NaN; +-Inf gives +-Inf or NaN depending on the signs (and zeros) of the dequantised weights it meets.  Rows
    without non-finite values must not be disturbed."""
    rng = np.random.default_rng(M * 7 + (force + 2))
    N, K = 300, 1024
    packed, scales, zps = _weights(rng, N, K)
    x = rng.standard_normal((M, K), dtype=np.float32)
    x[0, 5] = np.nan                                     # row 0: NaN
    if M > 1:
        x[1, 17] = np.inf                                # row 1: +Inf (meets weights of both signs and zeros)
        x[1, 700] = -np.inf
    if M > 4:
        x[4, 3] = -np.inf                                # row 4: a single -Inf
    X = cuda(x).to(dtype)
    xr = X.float().cpu().numpy()
    with np.errstate(all="ignore"):
        ref = oracle.reference_quantized_linear(xr, packed, scales, zps)         # fp32, the reference's order
    pkg._lib.tune("force_path", force)
    try:
        y = pkg._lib.linear_fwd(X, cuda(packed), cuda(scales), cuda(zps), out_dtype=torch.float32).cpu().numpy()
    finally:
        pkg._lib.tune("force_path", -1)
    _same_nonfinite(y, ref)
    clean_rows = [m for m in range(M) if np.isfinite(xr[m]).all()]
    assert np.isfinite(ref[clean_rows]).all() and np.isfinite(y[clean_rows]).all()
    ref64 = oracle.reference_quantized_linear(np.where(np.isfinite(xr), xr, 0.0), packed, scales, zps, acc=np.float64)
    clean = [m for m in range(M) if np.isfinite(xr[m]).all()]
    if clean:
        assert np.abs(y[clean] - ref64[clean]).max() <= 1e-4 * np.abs(ref64[clean]).max()


def test_nonfinite_same_result_for_every_batch_size(oracle, pkg):
    """The same row gives the same non-finite pattern whether it is computed alone (M = 1), in a decode batch
    (M = 16) or in a prefill batch (M = 64)."""
    rng = np.random.default_rng(3)
    N, K = 256, 2048
    packed, scales, zps = _weights(rng, N, K)
    row = rng.standard_normal(K).astype(np.float32)
    row[100] = np.inf
    row[101] = np.nan
    P, S, Z = cuda(packed), cuda(scales), cuda(zps)
    outs = []
    for M in (1, 16, 64):
        x = rng.standard_normal((M, K), dtype=np.float32)
        x[M - 1] = row
        outs.append(pkg._lib.linear_fwd(cuda(x), P, S, Z).cpu().numpy()[M - 1])
    for o in outs:
        assert np.isnan(o).all()


@pytest.mark.parametrize("M", [1, 4, 16, 40])
@pytest.mark.parametrize("N,K", [(300, 1024), (64, 90)])
def test_bias(oracle, pkg, M, N, K):
    """from_linear with a bias (python/module.py:84 asserts there is none)."""
    torch.manual_seed(N + M)
    lin = torch.nn.Linear(K, N, bias=True)
    ql = pkg.QuantizedLinear.from_linear(lin.cuda())
    assert ql.bias is not None and "bias" in ql.state_dict()
    p, s, z = oracle.quantize_weights(lin.weight.detach().cpu().numpy())
    x = torch.randn(M, K)
    ref = oracle.reference_quantized_linear(x.numpy(), p, s, z, acc=np.float64) + lin.bias.detach().cpu().numpy().astype(np.float64)
    y = ql(x.cuda()).cpu().numpy()
    assert np.abs(y - ref).max() <= 2e-5 * np.abs(ref).max() + 1e-6
    ql2 = pkg.QuantizedLinear(K, N, bias=True).cuda()
    ql2.load_state_dict(ql.state_dict())
    assert torch.equal(ql2(x.cuda()), ql(x.cuda()))
    # a module without bias keeps exactly the reference's three state_dict keys
    assert sorted(pkg.QuantizedLinear(K, N).state_dict()) == ["packed_weights", "scales", "zero_points"]


@pytest.mark.parametrize("K", [256, 90])                 # tcgen05 grouped path / generic path
def test_grouped_fwd_zero_fills_rows_outside_every_group(oracle, pkg, K):
    """include/b200q.h: 'Rows outside every group are zero-filled' (csrc/moe_int4_kernel.cu:109 zero-initialises
    the output) -- also when offsets[0] > 0 or offsets[E] < R, on both kernels behind the entry point."""
    rng = np.random.default_rng(K)
    E, N, R = 3, 128, 70
    packed = rng.integers(0, 256, size=(E, N, K // 2), dtype=np.uint8)
    scales = (rng.random((E, N), dtype=np.float32) * 0.01 + 0.001).astype(np.float32)
    zps = rng.integers(0, 16, size=(E, N)).astype(np.float32)
    xs = rng.standard_normal((R, K), dtype=np.float32)
    offsets = np.array([4, 20, 20, 61], dtype=np.int32)              # rows 0..3 and 61..69 belong to no expert
    lib = pkg._lib.load()
    X, P, S, Z, O = cuda(xs), cuda(packed), cuda(scales), cuda(zps), cuda(offsets)
    y = torch.full((R, N), 7.0, device="cuda")                       # poison: every row must be written
    nb = lib.b200q_moe_grouped_ws_bytes(R, E, N, K)
    ws = torch.zeros(max(nb, 16), dtype=torch.uint8, device="cuda")
    pkg._lib.check(lib.b200q_moe_grouped_fwd(X.data_ptr(), 0, P.data_ptr(), S.data_ptr(), Z.data_ptr(), O.data_ptr(), E,
                                             y.data_ptr(), 0, R, N, K, ws.data_ptr(), ws.numel(),
                                             torch.cuda.current_stream().cuda_stream), "grouped")
    y = y.cpu().numpy()
    assert not y[:4].any() and not y[61:].any()
    for e in range(E):
        lo, hi = offsets[e], offsets[e + 1]
        if hi > lo:
            ref = oracle.reference_quantized_linear(xs[lo:hi], packed[e], scales[e], zps[e], acc=np.float64)
            assert np.abs(y[lo:hi] - ref).max() <= 1e-4 * np.abs(ref).max()


def test_router_nan_logits_stay_in_range_and_propagate(oracle, pkg):
    """A NaN logit: torch.softmax makes the token's probabilities NaN and the NaN reaches the output through the
    routing weights (routing.py:72-76, 186-187).  The expert indices must stay valid (they index the permutation and
    the gather), and the other tokens must be untouched."""
    rng = np.random.default_rng(0)
    T, E, k, d, F = 50, 8, 2, 128, 256
    logits = rng.standard_normal((T, E), dtype=np.float32)
    logits[7, 3] = np.nan
    logits[9, 0] = np.inf
    dr = pkg.route(cuda(logits), k)
    idx = dr.expert_indices.cpu().numpy()
    assert idx.min() >= 0 and idx.max() < E
    assert all(len(set(r)) == k for r in idx)
    w = dr.expert_weights.cpu().numpy()
    assert np.isnan(w[7]).all() and np.isnan(w[9]).all()
    clean = [t for t in range(T) if t not in (7, 9)]
    idx0, w0 = oracle.softmax_topk(logits[clean], k)
    assert np.array_equal(idx[clean], idx0) and np.allclose(w[clean], w0, rtol=1e-6)
    assert int(dr.counts.sum()) == T * k and int(dr.offsets[-1]) == T * k
    inv = dr.inv_perm.cpu().numpy()
    assert sorted(inv.tolist()) == list(range(T * k))
    ws = [torch.from_numpy((rng.standard_normal((F, d)) * 0.05).astype(np.float16)).cuda() for _ in range(E)]
    moe = pkg.QuantizedMoE.from_fp16_weights(ws)
    out = moe.forward_routed(cuda(rng.standard_normal((T, d), dtype=np.float32)), cuda(logits), top_k=k).cpu().numpy()
    assert np.isnan(out[7]).all() and np.isnan(out[9]).all() and np.isfinite(out[clean]).all()


def test_permute_with_out_of_range_expert_ids_is_memory_safe(pkg):
    """Caller-made routing with a bad expert id: the assignment is dropped, every output is defined."""
    T, E, k = 40, 4, 2
    idx = torch.randint(0, E, (T, k), dtype=torch.int32)
    idx[3, 1] = 99
    idx[10, 0] = -5
    counts, offsets, sorted_slot, inv_perm = pkg._lib.moe_permute(idx.cuda(), E)
    assert int(counts.sum()) == T * k - 2 and int(offsets[-1]) == T * k - 2
    ss, inv = sorted_slot.cpu().numpy(), inv_perm.cpu().numpy()
    assert (ss[T * k - 2:] == -1).all() and inv[3 * k + 1] == -1 and inv[10 * k] == -1
    x = torch.randn(T, 64, device="cuda")
    xs = pkg._lib.moe_gather_rows(x, sorted_slot, k)
    assert not xs[T * k - 2:].any()
    w = torch.full((T, k), 0.5, device="cuda")
    out = pkg._lib.moe_combine(xs, inv_perm, w, k).cpu()
    assert torch.isfinite(out).all()
    assert torch.allclose(out[3], 0.5 * x[3].cpu())                # one of token 3's two slots was dropped


def test_stacked_weight_cache_follows_load_state_dict_and_expert_replacement(oracle, pkg):
    """load_state_dict / expert replacement after a first forward must not leave stale stacked weights."""
    rng = np.random.default_rng(1)
    E, d, F, T = 4, 128, 256, 24
    mk = lambda: [torch.from_numpy((rng.standard_normal((F, d)) * 0.05).astype(np.float16)).cuda() for _ in range(E)]
    a, b = pkg.QuantizedMoE.from_fp16_weights(mk()), pkg.QuantizedMoE.from_fp16_weights(mk())
    x = cuda(rng.standard_normal((T, d), dtype=np.float32))
    logits = cuda(rng.standard_normal((T, E), dtype=np.float32))
    ya, yb = a.forward_routed(x, logits), b.forward_routed(x, logits)
    assert not torch.equal(ya, yb)
    a.load_state_dict(b.state_dict())
    assert torch.equal(a.forward_routed(x, logits), yb)
    c = pkg.QuantizedMoE.from_fp16_weights(mk())
    yc = c.forward_routed(x, logits)
    for i in range(E):
        b.experts[i] = c.experts[i]
    assert torch.equal(b.forward_routed(x, logits), yc)


# ---- fused gate + up pair; 8(f)4: group-wise scales ---------------------------------------------------
@pytest.mark.parametrize("M", [1, 2, 5, 16, 40, 300])
@pytest.mark.parametrize("d,F", [(256, 512), (1024, 2816), (4096, 11008)])
def test_fused_gate_up_mlp(oracle, pkg, M, d, F):
    """QuantizedGatedMLP = down(silu(gate(x)) * up(x)) in two launches, against the float64 oracle composed from the
    reference's primitives, and against the three layers run one by one; state_dict compatible with three reference
    QuantizedLinear modules."""
    if (d, F) == (4096, 11008) and M not in (1, 16, 300):
        pytest.skip("full Llama size: three batch sizes are enough")
    torch.manual_seed(d + F)
    gate, up, down = (torch.nn.Linear(d, F, bias=False), torch.nn.Linear(d, F, bias=False), torch.nn.Linear(F, d, bias=False))
    mlp = pkg.QuantizedGatedMLP.from_linears(gate.cuda(), up.cuda(), down.cuda())
    x = torch.randn(M, d)
    q = lambda lin: oracle.quantize_weights(lin.weight.detach().cpu().numpy())
    qg, qu, qd = q(gate), q(up), q(down)
    g = oracle.reference_quantized_linear(x.numpy(), *qg, acc=np.float64)
    u = oracle.reference_quantized_linear(x.numpy(), *qu, acc=np.float64)
    h = (oracle.silu(g) * u).astype(np.float32)                     # the product keeps h in fp32 (the activations' dtype)
    ref = oracle.reference_quantized_linear(h, *qd, acc=np.float64)
    y = mlp(x.cuda()).cpu().numpy()
    assert np.abs(y - ref).max() <= 2e-4 * np.abs(ref).max() + 1e-6
    # the sub-modules still work on their own (strided views of the interleaved stack) and carry the reference's keys
    y3 = mlp.down_proj(torch.nn.functional.silu(mlp.gate_proj(x.cuda())) * mlp.up_proj(x.cuda())).cpu().numpy()
    assert np.abs(y3 - ref).max() <= 2e-4 * np.abs(ref).max() + 1e-6
    assert sorted(mlp.state_dict()) == sorted(f"{m}.{b}" for m in ("gate_proj", "up_proj", "down_proj")
                                                for b in ("packed_weights", "scales", "zero_points"))
    assert np.array_equal(mlp.gate_proj.packed_weights.cpu().numpy(), qg[0]) and np.array_equal(mlp.up_proj.scales.cpu().numpy(), qu[1])
    mlp2 = pkg.QuantizedGatedMLP(d, F).cuda()
    mlp2.load_state_dict(mlp.state_dict())
    assert torch.equal(mlp2(x.cuda()), mlp(x.cuda()))


def test_fused_gate_up_nonfinite(oracle, pkg):
    """NaN / Inf rows of x through the gated decode kernel: same pattern as the three separate layers."""
    torch.manual_seed(0)
    d, F = 512, 1024
    mlp = pkg.QuantizedGatedMLP.from_linears(torch.nn.Linear(d, F, bias=False).cuda(), torch.nn.Linear(d, F, bias=False).cuda(),
                                             torch.nn.Linear(F, d, bias=False).cuda())
    x = torch.randn(4, d)
    x[1, 3] = float("nan")
    x[2, 7] = float("inf")
    y = mlp(x.cuda()).cpu().numpy()
    y3 = mlp.down_proj(torch.nn.functional.silu(mlp.gate_proj(x.cuda())) * mlp.up_proj(x.cuda())).cpu().numpy()
    assert np.isnan(y[1]).all() and np.array_equal(np.isnan(y), np.isnan(y3)) and np.isfinite(y[[0, 3]]).all()


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("M,N,K,G", [(1, 11008, 4096, 128), (4, 11008, 4096, 128), (16, 4096, 11008, 256), (9, 300, 1024, 128),
                                     (27, 4096, 4096, 128), (5, 14336, 4096, 128)])
def test_groupwise_scales_on_the_decode_kernel(oracle, pkg, dtype, M, N, K, G):
    """Decode-sized batches with groups of 128 / 256 columns run on the mid-batch decode kernel (gemv_hm.cu, GS instances:
    the two 128-column halves of every 256-column pair meet their own scale and zero point): against the float64 oracle
    at full Llama / Mixtral size, against the reference-speed SIMT kernel (force_path 1), NaN / Inf rows, determinism."""
    rng = np.random.default_rng(M + N + K + G)
    packed = rng.integers(0, 256, size=(N, K // 2), dtype=np.uint8)
    scales = (rng.random((N, K // G), dtype=np.float32) * 0.01 + 0.001).astype(np.float32)
    zps = rng.integers(0, 16, size=(N, K // G)).astype(np.float32)
    x = rng.standard_normal((M, K), dtype=np.float32)
    X = cuda(x).to(dtype)
    xr = X.float().cpu().numpy()
    P, S, Z = cuda(packed), cuda(scales), cuda(zps)
    rows = np.arange(N) if N <= 512 else rng.choice(N, size=256, replace=False)
    q = oracle.unpack_nibbles(packed[rows]).astype(np.float64)
    wd = (q - np.repeat(zps[rows], G, axis=1)) * np.repeat(scales[rows], G, axis=1)
    ref = xr.astype(np.float64) @ wd.T
    y = pkg._lib.linear_groupwise_fwd(X, P, S, Z, G, out_dtype=torch.float32).cpu().numpy()
    assert np.abs(y[:, rows] - ref).max() <= 4e-6 * np.abs(ref).max()
    assert np.array_equal(pkg._lib.linear_groupwise_fwd(X, P, S, Z, G, out_dtype=torch.float32).cpu().numpy(), y)
    pkg._lib.tune("force_path", 1)
    try:
        yg = pkg._lib.linear_groupwise_fwd(X, P, S, Z, G, out_dtype=torch.float32).cpu().numpy()
    finally:
        pkg._lib.tune("force_path", -1)
    assert np.abs(y - yg).max() <= 1e-4 * np.abs(yg).max()
    if M >= 4:
        xn = xr.copy()
        xn[1, 7] = np.nan
        xn[2, K - 3] = np.inf
        yn = pkg._lib.linear_groupwise_fwd(cuda(xn).to(dtype), P, S, Z, G, out_dtype=torch.float32).cpu().numpy()
        with np.errstate(all="ignore"):
            refn = (xn.astype(np.float32) @ wd.astype(np.float32).T)
        assert np.isnan(yn[1]).all() and np.array_equal(np.isnan(yn[2][rows]), np.isnan(refn[2]))
        ok = [m for m in range(M) if m not in (1, 2)]
        assert np.abs(yn[ok][:, rows] - ref[ok]).max() <= 4e-6 * np.abs(ref).max()


@pytest.mark.parametrize("G", [32, 128, 256, 512])
@pytest.mark.parametrize("M", [1, 8, 20, 40])
def test_groupwise_scales(oracle, pkg, G, M):
    """One scale / zero point per G columns: packing / scales / zero points bit-exact with the reference's formulas
    applied per group (the oracle's quantize_weights on W viewed as [N K/G, G]); forward against the float64 oracle;
    finer groups follow an outlier-heavy weight matrix better than per-row scales."""
    rng = np.random.default_rng(G + M)
    N, K = 192, 1024
    w = rng.standard_normal((N, K)).astype(np.float32)
    w[:, ::97] *= 20.0                                              # outliers: what group-wise scales are for
    p0, s0, z0 = oracle.quantize_weights(w.reshape(N * (K // G), G))
    p, s, z = pkg.quantize_weights_grouped(torch.from_numpy(w).cuda(), G)
    assert np.array_equal(p.cpu().numpy(), p0.reshape(N, K // 2))
    assert np.array_equal(s.cpu().numpy(), s0.reshape(N, K // G)) and np.array_equal(z.cpu().numpy(), z0.reshape(N, K // G))
    wd = oracle.dequantize_weights(p0, s0, z0).reshape(N, K)
    assert np.array_equal(pkg.dequantize_weights_grouped(p, s, z).cpu().numpy(), wd)
    lin = torch.nn.Linear(K, N, bias=True)
    lin.weight.data = torch.from_numpy(w)
    ql = pkg.QuantizedLinear.from_linear(lin.cuda(), group_size=G)
    assert ql.scales.shape == (N, K // G) and "group_size" in repr(ql)
    x = rng.standard_normal((M, K)).astype(np.float32)
    ref = x.astype(np.float64) @ wd.astype(np.float64).T + lin.bias.detach().cpu().numpy().astype(np.float64)
    y = ql(torch.from_numpy(x).cuda()).cpu().numpy()
    assert np.abs(y - ref).max() <= 1e-4 * np.abs(ref).max()
    ql2 = pkg.QuantizedLinear(K, N, bias=True, group_size=G).cuda()
    ql2.load_state_dict(ql.state_dict())
    assert torch.equal(ql2(torch.from_numpy(x).cuda()), ql(torch.from_numpy(x).cuda()))
    # accuracy: group-wise error is well below the per-row error on this matrix
    wr = oracle.dequantize_weights(*oracle.quantize_weights(w))
    if G <= 64:                                          # (an outlier every 97 columns: groups of >= 128 columns hold one too)
        assert np.abs(wd - w).mean() < 0.5 * np.abs(wr - w).mean()
