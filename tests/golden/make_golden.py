"""Generate tests/golden/*.npz by running the REAL reference (mounted at /root/reference) on CPU.

Run in the build container only (the reference does not travel to the GPU box):

    python tests/golden/make_golden.py

The fixtures hold the reference's outputs (and, where the input is not a torch-seeded tensor that the
tests can regenerate, the inputs too).  Large outputs are stored as sha256 digests.
"""
import hashlib
import os
import sys

import numpy as np
import torch

REF = os.environ.get("B200Q_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
from python.quantize import quantize_weights, dequantize_weights, reference_quantized_linear  # noqa: E402
from python.module import QuantizedLinear  # noqa: E402
from benchmark.moe_grouped_gemm import routing as ref_routing  # noqa: E402
from benchmark.moe_grouped_gemm.moe_int4_module import QuantizedMoE  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
torch.set_num_threads(8)


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def npy(t: torch.Tensor) -> np.ndarray:
    return t.detach().cpu().numpy()


def quantize_cases():
    out = {}
    # the reference's own round-trip tests: tests/test_correctness.py:49 (16x32, seed 42),
    # :59 (256x512, seed 123), :201 (64x128, seed 42), :221 (256x512, seed 42)
    for name, seed, shape in (("s42_16x32", 42, (16, 32)), ("s123_256x512", 123, (256, 512)),
                              ("s42_64x128", 42, (64, 128)), ("s42_256x512", 42, (256, 512))):
        torch.manual_seed(seed)
        w = torch.randn(*shape)
        p, s, z = quantize_weights(w)
        out[f"{name}_w"] = npy(w)
        out[f"{name}_packed"], out[f"{name}_scales"], out[f"{name}_zp"] = npy(p), npy(s), npy(z)
        out[f"{name}_deq"] = npy(dequantize_weights(p, s, z))
    # tests/test_correctness.py:68 / :236 (4096x4096, seed 7) -- digests only
    torch.manual_seed(7)
    w = torch.randn(4096, 4096)
    x = torch.randn(4096)
    p, s, z = quantize_weights(w)
    out["s7_4096_packed_sha"] = np.array(sha(npy(p)))
    out["s7_4096_scales"], out["s7_4096_zp"] = npy(s), npy(z)
    out["s7_4096_packed_sum"] = np.array(int(npy(p).astype(np.int64).sum()))
    out["s7_4096_y"] = npy(reference_quantized_linear(x, p, s, z))       # test_cuda_large_dims oracle output
    # edge cases: constant rows (tests/test_correctness.py:93), negative constant, zero row, tiny range,
    # exact .5 rounding ties, wide dynamic range
    edge = torch.stack([
        torch.ones(8) * 3, -torch.ones(8) * 3, torch.zeros(8), torch.ones(8) * 1e-12,
        torch.tensor([0.0, 0.5, 1.0, 1.5, 2.0, 2.5, 3.0, 7.5]),
        torch.tensor([-7.5, -2.5, -1.5, -0.5, 0.5, 1.5, 2.5, 7.5]),
        torch.tensor([1e30, -1e30, 0.0, 1.0, -1.0, 5e29, -5e29, 1e-30]),
        torch.tensor([1e-40, 2e-40, 3e-40, 0.0, -1e-40, 5e-41, 1e-39, -1e-39]),
    ])
    p, s, z = quantize_weights(edge)
    out["edge_w"], out["edge_packed"], out["edge_scales"], out["edge_zp"] = npy(edge), npy(p), npy(s), npy(z)
    out["edge_deq"] = npy(dequantize_weights(p, s, z))
    # from_linear on the BASELINE config (seed 42, nn.Linear(4096, 11008)): digest + ranges
    torch.manual_seed(42)
    lin = torch.nn.Linear(4096, 11008, bias=False)
    ql = QuantizedLinear.from_linear(lin)
    out["llama_up_packed_sha"] = np.array(sha(npy(ql.packed_weights)))
    out["llama_up_scales_sha"] = np.array(sha(npy(ql.scales)))
    out["llama_up_zp_sha"] = np.array(sha(npy(ql.zero_points)))
    xx = torch.randn(4096)
    out["llama_up_y"] = npy(ql(xx))
    return out


def linear_cases():
    out = {}
    # tests/test_correctness.py:201-219
    torch.manual_seed(42)
    w = torch.randn(64, 128); x = torch.randn(128)
    p, s, z = quantize_weights(w)
    out["t1d_x"], out["t1d_y"] = npy(x), npy(reference_quantized_linear(x, p, s, z))
    # tests/test_correctness.py:221-234
    torch.manual_seed(42)
    w = torch.randn(256, 512); x = torch.randn(4, 512)
    p, s, z = quantize_weights(w)
    out["tb_x"], out["tb_y"] = npy(x), npy(reference_quantized_linear(x, p, s, z))
    # odd sizes for the generic path: K not a multiple of 8, N not a multiple of anything
    torch.manual_seed(5)
    w = torch.randn(37, 90); x = torch.randn(5, 90)
    p, s, z = quantize_weights(w)
    out["odd_w"], out["odd_x"] = npy(w), npy(x)
    out["odd_packed"], out["odd_scales"], out["odd_zp"] = npy(p), npy(s), npy(z)
    out["odd_y"] = npy(reference_quantized_linear(x, p, s, z))
    return out


def routing_cases():
    out = {}
    for dist in ("random", "skewed"):
        r = ref_routing.simulate_routing(64, 8, 2, distribution=dist, device="cpu", seed=42)
        # regenerate the logits the same way simulate_routing does (routing.py:51-69)
        torch.manual_seed(42)
        if dist == "random":
            logits = torch.randn(64, 8)
        else:
            ep = 1.0 / (torch.arange(8, dtype=torch.float32) + 1)
            ep = ep / ep.sum()
            logits = torch.log(ep + 1e-10).unsqueeze(0).expand(64, -1)
            logits = logits + torch.randn_like(logits) * 0.5
        out[f"{dist}64_logits"] = npy(logits)
        out[f"{dist}64_idx"] = npy(r.expert_indices)
        out[f"{dist}64_w"] = npy(r.expert_weights)
        out[f"{dist}64_counts"] = np.array(r.tokens_per_expert)
        out[f"{dist}64_offsets"] = np.array(r.expert_token_offsets)
        # dispatch + combine through the reference (routing.py:96-189) with identity "experts"
        torch.manual_seed(1)
        x = torch.randn(64, 16)
        xs, perm = ref_routing.create_expert_inputs(x, r, 8, 2)
        comb = ref_routing.combine_expert_outputs(xs, r, perm, 2)
        out[f"{dist}64_x"] = npy(x)
        out[f"{dist}64_combined_identity"] = npy(comb)
        xn = npy(x)
        for e in range(8):
            # the reference's argsort is unstable: only the SET of tokens per expert is defined
            toks = sorted(int(np.where((xn == row).all(axis=1))[0][0]) for row in npy(xs[e]))
            out[f"{dist}64_expert{e}_tokens"] = np.array(toks, dtype=np.int64)
        for T in (8192, 16384):
            rr = ref_routing.simulate_routing(T, 8, 2, distribution=dist, device="cpu", seed=42)
            out[f"{dist}{T}_counts"] = np.array(rr.tokens_per_expert)
            out[f"{dist}{T}_offsets"] = np.array(rr.expert_token_offsets)
            out[f"{dist}{T}_idx_sha"] = np.array(sha(npy(rr.expert_indices)))
            out[f"{dist}{T}_w_sum"] = np.array(float(rr.expert_weights.double().sum()))
    return out


def moe_cases():
    """QuantizedMoE (benchmark/moe_grouped_gemm/moe_int4_module.py) chained with routing.py on a
    debug-size layer (E=4, d=128, F=256, T=48, k=2), fp32 activations, plus the fp16 variant."""
    out = {}
    E, d, F, T, k = 4, 128, 256, 48, 2
    torch.manual_seed(3)
    ws = [torch.randn(F, d).half() * 0.02 for _ in range(E)]          # moe_int4_module.py:151-154 distribution
    x = torch.randn(T, d)
    logits = torch.randn(T, E)
    moe = QuantizedMoE.from_fp16_weights(ws)
    out["w_fp16"] = np.stack([npy(w) for w in ws])
    out["x"], out["logits"] = npy(x), npy(logits)
    out["packed"] = np.stack([npy(e.packed_weights) for e in moe.experts])
    out["scales"] = np.stack([npy(e.scales) for e in moe.experts])
    out["zp"] = np.stack([npy(e.zero_points) for e in moe.experts])
    # routing exactly as simulate_routing does after the logits (routing.py:72-86)
    rw = torch.softmax(logits, dim=-1)
    ew, ei = torch.topk(rw, k, dim=-1)
    ew = ew / ew.sum(dim=-1, keepdim=True)
    counts = [0] * E
    for i in ei.flatten().numpy():
        counts[i] += 1
    offs = [0]
    for c in counts[:-1]:
        offs.append(offs[-1] + c)
    r = ref_routing.RoutingResult(ei, ew, counts, offs)
    xs, perm = ref_routing.create_expert_inputs(x, r, E, k)
    ys = moe(xs)
    ys = [y.float() for y in ys]
    out["idx"], out["w"], out["counts"] = npy(ei), npy(ew), np.array(counts)
    out["y_single"] = npy(ref_routing.combine_expert_outputs(ys, r, perm, k))
    # fp16 activations (what the reference benches feed): per-expert outputs are fp16
    xs16, perm16 = ref_routing.create_expert_inputs(x.half(), r, E, k)
    ys16 = moe(xs16)
    out["y_single_fp16"] = npy(ref_routing.combine_expert_outputs(ys16, r, perm16, k))
    # gated layer composed from reference primitives
    torch.manual_seed(4)
    w1 = [torch.randn(F, d).half() * 0.02 for _ in range(E)]
    w3 = [torch.randn(F, d).half() * 0.02 for _ in range(E)]
    w2 = [torch.randn(d, F).half() * 0.02 for _ in range(E)]
    q1 = [quantize_weights(w.float()) for w in w1]
    q3 = [quantize_weights(w.float()) for w in w3]
    q2 = [quantize_weights(w.float()) for w in w2]
    ysg = []
    for e in range(E):
        xe = xs[e]
        g = xe @ dequantize_weights(*q1[e]).T
        u = xe @ dequantize_weights(*q3[e]).T
        h = torch.nn.functional.silu(g) * u
        ysg.append(h @ dequantize_weights(*q2[e]).T)
    out["g_w1"], out["g_w3"], out["g_w2"] = (np.stack([npy(w) for w in ww]) for ww in (w1, w3, w2))
    out["y_gated"] = npy(ref_routing.combine_expert_outputs(ysg, r, perm, k))
    return out


def moe_int4_module_cases():
    """python/moe_int4_module.py quantize_weights_moe (per-expert scalar scale) on CPU tensors."""
    sys.modules.pop("moe_int4_cuda", None)
    import importlib
    import io
    import contextlib
    with contextlib.redirect_stdout(io.StringIO()):
        m = importlib.import_module("python.moe_int4_module")
    out = {}
    torch.manual_seed(11)
    ws = [torch.randn(24, 32).half() * (0.02 * (e + 1)) for e in range(3)]
    p, s, z = m.quantize_weights_moe(ws)
    out["w_fp16"] = np.stack([npy(w) for w in ws])
    out["packed"], out["scales"], out["zp"] = npy(p), npy(s), npy(z)
    return out


if __name__ == "__main__":
    for name, fn in (("quantize", quantize_cases), ("linear", linear_cases), ("routing", routing_cases),
                     ("moe", moe_cases), ("moe_int4_module", moe_int4_module_cases)):
        data = fn()
        path = os.path.join(HERE, f"{name}.npz")
        np.savez_compressed(path, **data)
        print(f"{path}: {len(data)} arrays, {os.path.getsize(path) / 1024:.1f} KiB")
    with open(os.path.join(HERE, "PROVENANCE.txt"), "w") as f:
        f.write(f"generated by tests/golden/make_golden.py from {REF}\n"
                f"torch {torch.__version__} (CPU, {torch.get_num_threads()} threads), numpy {np.__version__}\n")
