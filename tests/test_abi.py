"""CPU-side checks of the drop-in boundary: libb200q.so loads, exports every symbol that
include/b200q.h declares, validates arguments before touching a device, and the Python surface
mirrors the reference's names and fails loudly without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import pytest
import torch

from conftest import ROOT


def header_symbols():
    text = open(os.path.join(ROOT, "include", "b200q.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200q_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(pkg):
    lib = pkg._lib.load()
    syms = header_symbols()
    assert len(syms) >= 20
    for name in syms:
        assert hasattr(lib, name), f"{name} declared in include/b200q.h but not exported"
    # and the ctypes table covers the header exactly
    assert sorted(pkg._lib.SIGNATURES) == syms


def test_version_and_error_string(pkg):
    lib = pkg._lib.load()
    assert lib.b200q_version() == 200
    assert isinstance(lib.b200q_last_error_string(), bytes)


def test_argument_validation_needs_no_device(pkg):
    lib = pkg._lib.load()
    EINVAL = -1
    # odd K
    assert lib.b200q_linear_fwd(None, 0, None, None, None, None, 0, 1, 8, 7, None, 0, 0, None) == EINVAL
    assert b"even K" in lib.b200q_last_error_string()
    # bad dtype
    assert lib.b200q_linear_fwd(None, 9, None, None, None, None, 0, 1, 8, 8, None, 0, 0, None) == EINVAL
    # null pointers with a non-empty problem
    assert lib.b200q_linear_fwd(None, 0, None, None, None, None, 0, 1, 8, 8, None, 0, 0, None) == EINVAL
    # empty problems are a no-op success (reference: torch::empty({0, N}))
    assert lib.b200q_linear_fwd(None, 0, None, None, None, None, 0, 0, 8, 8, None, 0, 0, None) == 0
    assert lib.b200q_quantize_rows(None, 4, 7, None, None, None, None) == EINVAL
    assert lib.b200q_moe_topk(None, 4, 300, 2, None, None, None) == EINVAL
    assert lib.b200q_moe_topk(None, 4, 8, 9, None, None, None) == EINVAL
    assert lib.b200q_tune_set(b"no_such_key", 1) == EINVAL
    assert lib.b200q_tune_set(b"gemv_pdl", -1) == 0


def test_workspace_sizes(pkg):
    lib = pkg._lib.load()
    assert lib.b200q_linear_ws_bytes(1, 11008, 4096) > 0        # split-K partials + tickets
    assert lib.b200q_moe_permute_ws_bytes(16384, 8, 2) >= 16 * 8 * 4
    assert lib.b200q_minmax_ws_bytes() >= 2 * 1024 * 4 + 4


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback(pkg):
    lib = pkg._lib.load()
    assert lib.b200q_device_check(-1) != 0
    x = torch.randn(4, 64)
    ql = pkg.QuantizedLinear(64, 32)
    with pytest.raises(RuntimeError):
        ql(x)                                   # reference would silently take the CPU path
    with pytest.raises(RuntimeError):
        pkg.quantize_weights(torch.randn(4, 8))
    import fused_quant_linear_cuda
    with pytest.raises(RuntimeError, match="input must be a CUDA tensor"):
        fused_quant_linear_cuda.forward(x, ql.packed_weights, ql.scales, ql.zero_points)


def test_module_surface_matches_reference(pkg):
    """Names / buffers / shapes / dtypes of python/module.py:59-64 and moe_int4_module.py:38-45."""
    ql = pkg.QuantizedLinear(128, 64)
    sd = ql.state_dict()
    assert list(sd) == ["packed_weights", "scales", "zero_points"]
    assert sd["packed_weights"].shape == (64, 64) and sd["packed_weights"].dtype == torch.uint8
    assert sd["scales"].shape == (64,) and sd["scales"].dtype == torch.float32
    assert sd["zero_points"].shape == (64,) and sd["zero_points"].dtype == torch.float32
    assert "bits=4" in repr(ql)
    # superset of python/module.py:84 (which asserts `bias is None`): an optional fp32 bias buffer
    qb = pkg.QuantizedLinear(128, 64, bias=True)
    assert list(qb.state_dict()) == ["packed_weights", "scales", "zero_points", "bias"] and qb.bias.shape == (64,)
    moe = pkg.QuantizedMoE(4, 64, 128)
    assert len(moe.experts) == 4 and moe.experts[0].packed_weights.shape == (128, 32)
    assert moe.total_memory_bytes == 4 * (128 * 32 + 128 * 4 * 2)
    m = pkg.MoEINT4(2, 64, 128)
    assert m.packed_weights.shape == (2, 128, 32) and m.scales.shape == (2, 128)
    import fused_quant_linear_cuda, moe_int4_cuda
    assert callable(fused_quant_linear_cuda.forward) and callable(moe_int4_cuda.forward)
    assert pkg.MIXTRAL_8x7B.num_experts == 8 and pkg.MIXTRAL_8x7B.ffn_dim == 14336
