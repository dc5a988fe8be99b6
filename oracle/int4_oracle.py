"""CPU oracle for the INT4 dequantize-linear / INT4 MoE hot path.   *** TEST INFRASTRUCTURE ***

A plain-numpy restatement of the reference's algorithm, function by function, with the reference
file:line each one follows.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import this module -- always as the checker or the
timed CPU baseline, never as part of the product path (the product has no CPU path at all).

Pinning: ``tests/test_oracle.py`` checks every function here against ``tests/golden/*.npz`` --
outputs of the REAL reference (``/root/reference/python/quantize.py``,
``benchmark/moe_grouped_gemm/routing.py`` and ``moe_int4_module.py``) produced by
``tests/golden/make_golden.py`` in the build container with torch 2.11.0 (CPU), and including the
reference's own test vectors (tests/test_correctness.py:49-103, 201-253).  Integer results
(packed bytes, zero points, routing indices, histograms, offsets) are pinned bit-exactly; fp32
results to the tolerance stated in each test.

All arithmetic is IEEE fp32 (numpy float32 ops are correctly rounded, as torch's CPU kernels are);
rounding is round-half-to-even (``np.rint`` == ``torch.round``).
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np

f32 = np.float32


# --------------------------------------------------------------------------- quantisation format
def quantize_weights(w: np.ndarray) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """python/quantize.py:38-124.  w [N,K] f32, K even -> packed u8 [N,K/2], scales f32 [N], zp f32 [N]."""
    w = np.ascontiguousarray(w, dtype=f32)
    assert w.ndim == 2 and w.shape[1] % 2 == 0
    max_val = f32(15.0)
    w_min = w.min(axis=1)                                           # :73
    w_max = w.max(axis=1)                                           # :74
    scales = (w_max - w_min) / max_val                              # :80
    constant = w_max == w_min                                       # :85
    const_scale = np.maximum(np.abs(w_max), f32(1.0)) / max_val     # :90
    safe = np.where(constant, const_scale, scales).astype(f32)      # :86-92
    safe = np.maximum(safe, f32(1e-8))                              # :94
    with np.errstate(all="ignore"):
        zp = np.rint((-w_min) / safe).astype(f32)                   # :100
    zp = _clamp(zp, f32(0.0), max_val)                              # :101
    q = np.rint(w / safe[:, None] + zp[:, None])                    # :106-108
    q = _clamp(q.astype(f32), f32(0.0), max_val).astype(np.uint8)   # :109
    even, odd = q[:, 0::2], q[:, 1::2]                              # :120-121
    packed = ((odd << 4) | even).astype(np.uint8)                   # :122
    return packed, safe, zp


def _clamp(v, lo, hi):
    """torch.clamp semantics incl. keeping -0.0 (min(max(v, lo), hi) with v returned on equality)."""
    out = np.where(v < lo, lo, v)
    return np.where(out > hi, hi, out).astype(f32)


def unpack_nibbles(packed: np.ndarray) -> np.ndarray:
    """python/quantize.py:152-163: low nibble -> even column, high nibble -> odd column."""
    n, kh = packed.shape
    q = np.empty((n, kh * 2), dtype=np.uint8)
    q[:, 0::2] = packed & 0x0F
    q[:, 1::2] = packed >> 4
    return q


def dequantize_weights(packed: np.ndarray, scales: np.ndarray, zp: np.ndarray) -> np.ndarray:
    """python/quantize.py:127-173: (q - zp) * scale, subtract then multiply, fp32."""
    q = unpack_nibbles(packed).astype(f32)
    return ((q - zp.astype(f32)[:, None]) * scales.astype(f32)[:, None]).astype(f32)


def reference_quantized_linear(x: np.ndarray, packed, scales, zp, acc=np.float32) -> np.ndarray:
    """python/quantize.py:176-202: F.linear(x, dequantize_weights(...)).  acc=np.float64 gives the
    error-attribution variant (exact fp32 weights, float64 accumulation)."""
    w = dequantize_weights(packed, scales, zp)
    return (x.astype(acc) @ w.astype(acc).T).astype(acc)


def quantize_weights_moe(weights: List[np.ndarray]):
    """python/moe_int4_module.py:19-80: ONE scale / zero point per expert from the expert's global
    min / max, broadcast to [E,F]; q = clamp(round(w/scale + zp), 0, 15); same nibble order (:63-76)."""
    E = len(weights)
    F, d = weights[0].shape
    packed = np.zeros((E, F, d // 2), dtype=np.uint8)
    scales = np.zeros((E, F), dtype=f32)
    zps = np.zeros((E, F), dtype=f32)
    for e, w in enumerate(weights):
        w32 = w.astype(f32)
        w_min, w_max = w32.min(), w32.max()                          # :46-47
        scale = f32((w_max - w_min) / f32(15.0))                     # :49 (fp32 tensor division)
        zp = float(np.rint(f32(-w_min) / scale))                     # :50 (python round == half-even)
        zp = f32(max(0.0, min(15.0, zp)))                            # :51
        scales[e] = scale
        zps[e] = zp
        q = np.clip(np.rint(w32 / scale + zp), 0, 15).astype(np.uint8)   # :57-59
        packed[e] = (q[:, 1::2] << 4) | q[:, 0::2]                   # :63-76
    return packed, scales, zps


# --------------------------------------------------------------------------------------- routing
def softmax_topk(logits: np.ndarray, k: int):
    """routing.py:72-76: softmax -> top-k (descending; ties -> lowest expert index, the rule this
    repo defines, torch.topk leaves ties unspecified) -> renormalise over the k winners."""
    x = logits.astype(f32)
    m = x.max(axis=-1, keepdims=True)
    e = np.exp(x - m).astype(f32)
    p = (e / e.sum(axis=-1, keepdims=True, dtype=f32)).astype(f32)
    idx = np.argsort(-p, axis=-1, kind="stable")[:, :k]
    w = np.take_along_axis(p, idx, axis=-1)
    w = (w / w.sum(axis=-1, keepdims=True, dtype=f32)).astype(f32)
    return idx.astype(np.int64), w


def histogram_offsets(idx: np.ndarray, num_experts: int):
    """routing.py:79-86: tokens per expert and exclusive offsets."""
    counts = np.bincount(idx.reshape(-1), minlength=num_experts).astype(np.int64)
    offsets = np.concatenate([[0], np.cumsum(counts)[:-1]]).astype(np.int64)
    return counts, offsets


def permutation(idx: np.ndarray):
    """routing.py:121-132 with the sort made STABLE (ascending flat assignment index inside an
    expert); returns (sorted_indices, inverse_perm)."""
    flat = idx.reshape(-1)
    sorted_indices = np.argsort(flat, kind="stable")
    inverse = np.argsort(sorted_indices, kind="stable")
    return sorted_indices.astype(np.int64), inverse.astype(np.int64)


def create_expert_inputs(x: np.ndarray, idx: np.ndarray, num_experts: int):
    """routing.py:96-149."""
    T, k = idx.shape
    sorted_indices, inverse = permutation(idx)
    token = sorted_indices // k                                      # flat_token_idx[sorted_indices]
    counts, _ = histogram_offsets(idx, num_experts)
    outs, off = [], 0
    for e in range(num_experts):
        outs.append(x[token[off:off + counts[e]]])
        off += counts[e]
    return outs, inverse


def combine_expert_outputs(expert_outputs: List[np.ndarray], weights: np.ndarray, inverse: np.ndarray, k: int):
    """routing.py:152-189: cat -> un-permute -> [T,k,F] * w[...,None] -> sum over k, fp32."""
    cat = np.concatenate(expert_outputs, axis=0)
    un = cat[inverse].astype(f32)
    T = un.shape[0] // k
    un = un.reshape(T, k, -1)
    prod = (un * weights.astype(f32)[:, :, None]).astype(f32)
    out = prod[:, 0]
    for s in range(1, k):
        out = (out + prod[:, s]).astype(f32)
    return out


# ------------------------------------------------------------------------------------- MoE layers
def expert_forward(x: np.ndarray, packed, scales, zp, acc=np.float32):
    """QuantizedMoEExpert.forward, benchmark/moe_grouped_gemm/moe_int4_module.py:63-72:
    x @ dequant(W).T (in x's dtype; here fp32 or the float64 attribution variant)."""
    if x.shape[0] == 0:
        return np.zeros((0, packed.shape[0]), dtype=acc)
    return reference_quantized_linear(x, packed, scales, zp, acc=acc)


def moe_single_projection(x, logits, experts, k: int, acc=np.float32):
    """simulate_routing -> create_expert_inputs -> QuantizedMoE.forward -> combine_expert_outputs
    (SURVEY.md 3(e)); experts = list of (packed, scales, zp)."""
    idx, w = softmax_topk(logits, k)
    xs, inverse = create_expert_inputs(x, idx, len(experts))
    ys = [expert_forward(xe, *ex, acc=acc).astype(f32) for xe, ex in zip(xs, experts)]
    return combine_expert_outputs(ys, w, inverse, k)


def silu(a):
    return a / (1.0 + np.exp(-a))


def moe_gated(x, logits, w1, w3, w2, k: int, acc=np.float32):
    """Gated MLP composed only from reference primitives:
    h = silu(xs @ deq(w1)^T) * (xs @ deq(w3)^T); y = h @ deq(w2)^T; combine with routing weights."""
    idx, w = softmax_topk(logits, k)
    xs, inverse = create_expert_inputs(x, idx, len(w1))
    ys = []
    for e, xe in enumerate(xs):
        g = expert_forward(xe, *w1[e], acc=acc)
        u = expert_forward(xe, *w3[e], acc=acc)
        h = (silu(g) * u).astype(acc)
        ys.append(expert_forward(h, *w2[e], acc=acc).astype(f32))
    return combine_expert_outputs(ys, w, inverse, k)


def moe_int4_forward(packed, scales, zps, inputs, tokens_per_expert, input_offsets, acc=np.float32):
    """Intended semantics of moe_int4_cuda.forward (csrc/moe_int4_kernel.cu:93-136): rows
    [input_offsets[e], +tokens_per_expert[e]) go through expert e; other rows stay zero (:109)."""
    T = inputs.shape[0]
    out = np.zeros((T, packed.shape[1]), dtype=acc)
    for e in range(packed.shape[0]):
        n, o = int(tokens_per_expert[e]), int(input_offsets[e])
        if n:
            out[o:o + n] = reference_quantized_linear(inputs[o:o + n], packed[e], scales[e], zps[e], acc=acc)
    return out
