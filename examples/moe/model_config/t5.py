"""T5-style MoE model configurations (reference ``examples/moe/model_config/t5.py:100-153``: d_model / d_ff / heads / layers per
size, every second feed-forward replaced by an 8-expert top-2 MoE layer with capacity factor 1.25)."""
from easyparallellibrary_b200.models.moe_transformer import MoEConfig


def t5_moe(size: str = "small", num_experts: int = 8, gating: str = "top2", capacity_factor: float = 1.25, **kw) -> MoEConfig:
  table = {
      "tiny": dict(d_model=64, d_ff=128, n_layer=2, n_head=4, vocab_size=512, n_positions=64),
      "small": dict(d_model=512, d_ff=2048, n_layer=6, n_head=8),
      "base": dict(d_model=768, d_ff=3072, n_layer=12, n_head=12),
      "large": dict(d_model=1024, d_ff=4096, n_layer=24, n_head=16),
  }
  cfg = dict(table[size], num_experts=num_experts, gating=gating, capacity_factor=capacity_factor)
  cfg.update(kw)
  return MoEConfig(**cfg)
