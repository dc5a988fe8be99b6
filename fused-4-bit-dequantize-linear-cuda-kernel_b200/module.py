"""QuantizedLinear: drop-in for the reference's python/module.py:33-138 on B200.

Same constructor, ``from_linear`` classmethod, buffer names / shapes / dtypes (so reference
``state_dict``s load), ``forward`` and ``extra_repr``.  ``forward`` always runs the fused sm_100a
kernel; a CPU tensor raises instead of silently taking a slow path.
Superset: any leading dims ``[..., K]``, fp16 / bf16 activations, an optional fp32 ``bias`` buffer (the reference
asserts ``linear.bias is None``, python/module.py:84; a module without bias has exactly the reference's state_dict).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _lib
from .quantize import quantize_weights, quantize_weights_grouped


class QuantizedLinear(nn.Module):
    _gen = 0    # bumped whenever a weight buffer is replaced or moved (QuantizedGatedMLP re-stacks lazily when it changes)

    def __setattr__(self, name, value):
        if name in ("packed_weights", "scales", "zero_points"):
            object.__setattr__(self, "_gen", self._gen + 1)
        super().__setattr__(name, value)

    def _apply(self, fn, *args, **kwargs):
        object.__setattr__(self, "_gen", self._gen + 1)
        return super()._apply(fn, *args, **kwargs)

    def __init__(self, in_features: int, out_features: int, bias: bool = False, group_size: int = 0):
        super().__init__()
        self.in_features = in_features
        self.out_features = out_features
        # group_size > 0: one scale / zero point per `group_size` columns (scales [N, K/G]).  0 = the
        # reference's per-row format
        self.group_size = group_size
        # python/module.py:59-64
        self.register_buffer("packed_weights",
                             torch.zeros(out_features, in_features // 2, dtype=torch.uint8))
        sshape = (out_features,) if not group_size else (out_features, in_features // group_size)
        self.register_buffer("scales", torch.zeros(sshape, dtype=torch.float32))
        self.register_buffer("zero_points", torch.zeros(sshape, dtype=torch.float32))
        if bias:
            self.register_buffer("bias", torch.zeros(out_features, dtype=torch.float32))
        else:
            self.bias = None
        # Weights written by a kernel that may still be running (from_linear) must not be
        # prefetched ahead of it; the first forward therefore runs without the static-weights flag.
        self._weights_settled = False
        object.__setattr__(self, "_next", None)   # not a sub-module: only a prefetch hint, see set_next

    def set_next(self, next_layer: "QuantizedLinear | None") -> "QuantizedLinear":
        """Tell this layer which fused linear follows it in the decode loop: its decode kernel then pulls
        that layer's packed weights into L2 behind its own weight stream, so HBM keeps streaming between
        the two launches.  A hint only (L2 is coherent): results are unchanged whatever runs next."""
        object.__setattr__(self, "_next", next_layer)
        return self

    @classmethod
    def from_linear(cls, linear: nn.Linear, group_size: int = 0) -> "QuantizedLinear":
        # python/module.py:84 asserts `linear.bias is None`; here the bias is kept in fp32 and added by the kernel
        module = cls(linear.in_features, linear.out_features, bias=linear.bias is not None, group_size=group_size)
        if linear.bias is not None:
            module.bias = linear.bias.data.detach().to(torch.float32).clone()
        if group_size:
            packed, scales, zp = quantize_weights_grouped(linear.weight.data, group_size)
        else:
            packed, scales, zp = quantize_weights(linear.weight.data)
        module.packed_weights = packed
        module.scales = scales
        module.zero_points = zp
        return module

    def _fast_state(self):
        """(packed, scales, zero_points, bias, device, plain) of the buffers as they are now.  nn.Module resolves a buffer
        name through __getattr__ (~0.4 us each, five per call: as much as the C call itself); this keeps the tensors in
        a tuple that is rebuilt whenever one of them has been replaced (identity check against _buffers)."""
        b = self._buffers
        st = self.__dict__.get("_fast")
        p = b["packed_weights"]
        bias = b.get("bias")
        if bias is None:
            bias = self.__dict__.get("bias")
        if st is None or st[0] is not p or st[1] is not b["scales"] or st[2] is not b["zero_points"] or st[3] is not bias:
            st = (p, b["scales"], b["zero_points"], bias, p.device, p.is_contiguous() and not self.group_size)
            self.__dict__["_fast"] = st
        return st

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if not x.is_cuda:
            if self.packed_weights.is_cuda and x.is_pinned():
                return self.forward_host(x)        # host-resident caller: copies + kernel enqueued by one C call
            raise RuntimeError("QuantizedLinear (b200) runs on CUDA only; move the module and the input to a B200 "
                               "(or pass a pinned host tensor to a module that lives on the GPU)")
        st = self._fast_state()
        if st[4] != x.device:
            raise RuntimeError(f"weights are on {st[4]}, input on {x.device}")
        ext = _lib.torch_ext()
        if ext is not None and st[5] and x.is_contiguous() and x.shape[-1] == self.in_features:
            # compiled binding: checks, output / workspace allocation, stream lookup and the C-ABI call in one C++ function
            nxt = self._next
            y = ext.linear_forward(x, st[0], st[1], st[2], st[3], None,
                                   _lib.FLAG_STATIC_WEIGHTS if self._weights_settled else _lib.FLAG_NONE,
                                   nxt._buffers["packed_weights"] if nxt is not None else None)
            self._weights_settled = True
            return y
        return self._forward_cuda(x)

    def _forward_cuda(self, x: torch.Tensor) -> torch.Tensor:
        lead = x.shape[:-1]
        if x.shape[-1] != self.in_features:
            raise RuntimeError(f"expected last dim {self.in_features}, got {x.shape[-1]}")
        if self.group_size:
            # group-wise scales: decode-sized batches with G % 128 == 0 on the mid-batch decode kernel, else the format's
            # reference-speed kernel; the bias is fused / added by the library
            x2 = x.reshape(-1, self.in_features).contiguous()
            y = _lib.linear_groupwise_fwd(x2, self.packed_weights, self.scales, self.zero_points, self.group_size,
                                          flags=_lib.FLAG_STATIC_WEIGHTS if self._weights_settled else _lib.FLAG_NONE,
                                          bias=self.bias, next_packed=self._next.packed_weights if self._next is not None else None)
            return y.reshape(*lead, self.out_features)
        if not self.packed_weights.is_contiguous():
            # buffers that are strided views of an interleaved gate / up stack (QuantizedGatedMLP): compact them for this call
            x2 = x.reshape(-1, self.in_features).contiguous()
            y = _lib.linear_fwd(x2, self.packed_weights.contiguous(), self.scales.contiguous(), self.zero_points.contiguous(), bias=self.bias)
            return y.reshape(*lead, self.out_features)
        ext = _lib.torch_ext()
        if ext is not None and x.is_contiguous():
            # compiled binding: checks, output / workspace allocation, stream lookup and the C-ABI call in one C++ function
            y = ext.linear_forward(x, self.packed_weights, self.scales, self.zero_points, self.bias, None,
                                   _lib.FLAG_STATIC_WEIGHTS if self._weights_settled else _lib.FLAG_NONE,
                                   self._next.packed_weights if self._next is not None else None)
            self._weights_settled = True
            return y
        x2 = x.reshape(-1, self.in_features)
        if not x2.is_contiguous():
            x2 = x2.contiguous()
        flags = _lib.FLAG_STATIC_WEIGHTS if self._weights_settled else _lib.FLAG_NONE
        nxt = self._next.packed_weights if self._next is not None else None
        y = _lib.linear_fwd(x2, self.packed_weights, self.scales, self.zero_points, flags=flags, next_packed=nxt,
                            bias=self.bias)
        self._weights_settled = True
        return y.reshape(*lead, self.out_features)

    def forward_host(self, x_host: torch.Tensor, out: torch.Tensor = None) -> torch.Tensor:
        """x in pinned host memory -> result in pinned host memory (asynchronous on the current stream of the
        module's device: synchronise the stream before reading `out`).  The arithmetic runs on the GPU."""
        x2 = x_host.reshape(-1, self.in_features)
        M = x2.shape[0]
        dev = self.packed_weights.device
        if self.bias is not None or self.group_size or not self.packed_weights.is_contiguous():
            # (bias / group-wise scales / strided views: the device path, with the two copies enqueued around it)
            y = self.forward(x2.to(dev, non_blocking=True))
            if out is None:
                out = torch.empty((M, self.out_features), dtype=x2.dtype).pin_memory()
            out.view(M, self.out_features).copy_(y, non_blocking=True)
            return out.reshape(*x_host.shape[:-1], self.out_features)
        st = getattr(self, "_stage", None)
        if st is None or st[0].shape[0] != M or st[0].dtype != x2.dtype:
            st = (torch.empty((M, self.in_features), dtype=x2.dtype, device=dev),
                  torch.empty((M, self.out_features), dtype=x2.dtype, device=dev))
            self._stage = st
        if out is None:
            out = torch.empty((M, self.out_features), dtype=x2.dtype).pin_memory()
        flags = _lib.FLAG_STATIC_WEIGHTS if self._weights_settled else _lib.FLAG_NONE
        ext = _lib.torch_ext()
        fs = self._fast_state()
        if ext is not None and fs[5] and x2.is_contiguous() and out.is_contiguous() and out.dtype == x2.dtype:
            ext.linear_forward_host(x2, st[0], fs[0], fs[1], fs[2], st[1], out.view(M, self.out_features), flags)
        else:
            _lib.linear_fwd_host(x2, st[0], self.packed_weights, self.scales, self.zero_points, st[1], out, flags=flags)
        self._weights_settled = True
        return out.reshape(*x_host.shape[:-1], self.out_features)

    def _load_from_state_dict(self, *args, **kwargs):
        self._weights_settled = False
        return super()._load_from_state_dict(*args, **kwargs)

    def extra_repr(self) -> str:
        return (f"in_features={self.in_features}, "
                f"out_features={self.out_features}, "
                f"bits=4" + (f", group_size={self.group_size}" if self.group_size else ""))


class QuantizedGatedMLP(nn.Module):
    """`down(silu(gate(x)) * up(x))` -- a Llama-style MLP -- in TWO launches: the gate and up
    projections are one fused dequantize-linear whose epilogue applies the SiLU-gate (b200q_linear_gated_fwd: decode
    kernel for M <= 16, tcgen05 GEMM above), then the down projection.

    `gate_proj`, `up_proj`, `down_proj` are QuantizedLinear modules with the reference's buffer names, so a state_dict of
    three reference QuantizedLinear layers loads unchanged; the rows of gate and up live ONCE on the device, interleaved
    (2f: gate, 2f+1: up), and the two modules' buffers are strided views of that stack."""

    def __init__(self, hidden_dim: int, ffn_dim: int):
        super().__init__()
        self.hidden_dim, self.ffn_dim = hidden_dim, ffn_dim
        self.gate_proj = QuantizedLinear(hidden_dim, ffn_dim)
        self.up_proj = QuantizedLinear(hidden_dim, ffn_dim)
        self.down_proj = QuantizedLinear(ffn_dim, hidden_dim)
        self._stack = None
        self._stamp = None

    @classmethod
    def from_linears(cls, gate: nn.Linear, up: nn.Linear, down: nn.Linear) -> "QuantizedGatedMLP":
        assert gate.bias is None and up.bias is None, "the fused gate / up pair takes no bias"
        mlp = cls(gate.in_features, gate.out_features)
        mlp.gate_proj = QuantizedLinear.from_linear(gate)
        mlp.up_proj = QuantizedLinear.from_linear(up)
        mlp.down_proj = QuantizedLinear.from_linear(down)
        return mlp

    def _views(self, st):
        out = []
        for j, name in enumerate(("packed_weights", "scales", "zero_points")):
            out.append((self.gate_proj, name, st[j][0::2]))
            out.append((self.up_proj, name, st[j][1::2]))
        return out

    def stacked(self):
        stamp = (id(self.gate_proj), self.gate_proj._gen, id(self.up_proj), self.up_proj._gen)
        if self._stack is None or self._stamp != stamp:
            st = self._stack
            ok = st is not None and all(m._buffers[n].data_ptr() == v.data_ptr() and m._buffers[n].stride() == v.stride()
                                        and m._buffers[n].device == v.device for m, n, v in self._views(st))
            if not ok:
                dev = self.gate_proj.packed_weights.device
                F, d = self.ffn_dim, self.hidden_dim
                st = (torch.empty((2 * F, d // 2), dtype=torch.uint8, device=dev),
                      torch.empty((2 * F,), dtype=torch.float32, device=dev), torch.empty((2 * F,), dtype=torch.float32, device=dev))
                for m, n, v in self._views(st):
                    v.copy_(m._buffers[n].to(dev))
                    m._buffers[n] = v
                self._stack = st
            self._stamp = stamp
        return self._stack

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        _lib.require_cuda(x, "x")
        lead = x.shape[:-1]
        x2 = x.reshape(-1, self.hidden_dim)
        if not x2.is_contiguous():
            x2 = x2.contiguous()
        p13, s13, z13 = self.stacked()
        h = _lib.linear_gated_fwd(x2, p13, s13, z13, flags=_lib.FLAG_STATIC_WEIGHTS, next_packed=self.down_proj.packed_weights)
        return self.down_proj(h).reshape(*lead, self.hidden_dim)


def link_decode_order(layers, cyclic: bool = True):
    """Wire the next-layer L2 prefetch hints of the fused linears of a decode loop, in the order they run:
    layers[i].set_next(layers[i + 1]); with `cyclic` the last one points at the first (the next token's first
    layer).  Returns `layers`."""
    layers = list(layers)
    for i, layer in enumerate(layers):
        if i + 1 < len(layers):
            layer.set_next(layers[i + 1])
        elif cyclic and len(layers) > 1:
            layer.set_next(layers[0])
    return layers
