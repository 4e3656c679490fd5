"""Human-readable summary of a parallel plan (reference: ``utils/summary_info.py`` + ``Graph.format()``,
``ir/graph.py:587-598``): which stage / split shard lives on which rank, replica count, micro-batches, bucket sizes."""
from __future__ import annotations

from typing import List


def plan_summary(trainer) -> str:
  plan = trainer.plan
  lines: List[str] = []
  lines.append("EPL plan: world=%d stages=%d replicas=%d micro_batches=%d schedule=%s" % (
      plan.world, plan.num_stages, plan.num_replicas, plan.num_micro_batch, trainer.config.pipeline.strategy))
  for s, reps in enumerate(plan.stage_ranks):
    lines.append("  stage %d -> ranks %s" % (s, reps))
  for key, flat in (getattr(trainer, "flats", None) or {}).items():
    buckets = getattr(flat, "buckets", [])
    n = sum(getattr(b, "numel", 0) for b in buckets) if buckets else 0
    lines.append("  param group %s: %d buckets, %.1f M elements%s" % (
        key, len(buckets), n / 1e6, " (fused reduce-scatter + Adam + all-gather)" if getattr(trainer, "fused", None) is not None else ""))
  cfg = trainer.config
  feats = []
  if cfg.zero.level:
    feats.append("zero=" + cfg.zero.level)
  if cfg.gradient_checkpoint.type:
    feats.append("gc=" + cfg.gradient_checkpoint.type)
  if cfg.amp.level:
    feats.append("amp=" + cfg.amp.level)
  if cfg.offload.level:
    feats.append("offload=" + cfg.offload.level)
  if feats:
    lines.append("  features: " + " ".join(feats))
  return "\n".join(lines)
