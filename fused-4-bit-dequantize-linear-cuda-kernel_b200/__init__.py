"""B200-native fused INT4 dequantize-linear and INT4 MoE hot path.

Drop-in surface of samy19980109/Fused-4-bit-Dequantize-Linear-CUDA-Kernel (python/__init__.py:14-22
and benchmark/moe_grouped_gemm): same names, same buffers, same results -- computed by hand-written
sm_100a kernels in ``libb200q.so`` (C ABI: include/b200q.h).
"""
from .quantize import (quantize_weights, dequantize_weights, reference_quantized_linear, quantize_weights_grouped,
                       dequantize_weights_grouped)
from .module import QuantizedLinear, QuantizedGatedMLP, link_decode_order
from .moe import QuantizedMoEExpert, QuantizedMoE, MoEINT4, quantize_weights_moe
from .routing import (RoutingResult, DeviceRouting, simulate_routing, create_expert_inputs,
                      combine_expert_outputs, get_expert_sizes_for_benchmark, route, make_logits)
from .config import MoEConfig, MIXTRAL_8x7B, DEBUG_CONFIG, LLAMA_7B_MLP
from .ep import ExpertParallelMoE, dispatch_plan, shard_experts, local_expert_list, virtual_ids, ep_plan_host
from . import _lib

__all__ = [
    "link_decode_order", "QuantizedGatedMLP", "quantize_weights_grouped", "dequantize_weights_grouped",
    "quantize_weights", "dequantize_weights", "reference_quantized_linear", "QuantizedLinear",
    "QuantizedMoEExpert", "QuantizedMoE", "MoEINT4", "quantize_weights_moe",
    "RoutingResult", "DeviceRouting", "simulate_routing", "create_expert_inputs",
    "combine_expert_outputs", "get_expert_sizes_for_benchmark", "route", "make_logits",
    "ExpertParallelMoE", "dispatch_plan", "shard_experts", "local_expert_list", "virtual_ids", "ep_plan_host",
    "MoEConfig", "MIXTRAL_8x7B", "DEBUG_CONFIG", "LLAMA_7B_MLP",
]
