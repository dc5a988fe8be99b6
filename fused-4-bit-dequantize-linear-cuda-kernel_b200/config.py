"""MoE shape configs (benchmark/moe_grouped_gemm/config.py:34-109) -- shapes only."""
from dataclasses import dataclass


@dataclass
class MoEConfig:
    name: str
    num_experts: int
    hidden_dim: int
    ffn_dim: int
    top_k: int

    @property
    def expert_params(self) -> int:
        return self.hidden_dim * self.ffn_dim


MIXTRAL_8x7B = MoEConfig("Mixtral-8x7B", 8, 4096, 14336, 2)        # config.py:70-76
DEBUG_CONFIG = MoEConfig("Debug", 4, 256, 512, 2)
LLAMA_7B_MLP = ((4096, 11008), (11008, 4096))                      # (in_features, out_features)
