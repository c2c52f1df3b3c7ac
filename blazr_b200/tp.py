"""Tensor-parallel shard planning for the quantized projections (host logic, no GPU needed).

Mirrors blazr's rule (reference src/engine/tensor_parallel.rs:61-67 shard_range: even split, remainder to the
low ranks; :76-101 validate_tp_config: heads and kv-heads must divide tp) applied at the granularity the packed
formats allow:
  * column-parallel (q, k, v by heads; gate, up by rows of 256): split N, no communication;
  * row-parallel (o by heads, down by 256-k super-blocks): split K at block granularity, all-reduce(sum) after;
  * lm_head: column-parallel over the vocabulary at 128-row granularity + all-gather.
GGUF is rejected for TP by the reference loader (src/loader/api.rs:52-54); here every format shards because
the split happens on packed blocks before the upload repack.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Tuple

from . import ops


@dataclass
class TpPlan:
    rank: int
    world: int
    n_heads: int          # local q heads
    n_kv_heads: int       # local kv heads
    q_rows: Tuple[int, int]
    kv_rows: Tuple[int, int]
    o_cols: Tuple[int, int]
    ffn_rows: Tuple[int, int]
    down_cols: Tuple[int, int]
    vocab_rows: Tuple[int, int]


def vocab_shard_rows(vocab: int, world: int) -> int:
    """rows of the logits buffer every rank contributes to the lm_head all-gather (128-row granularity)"""
    per = -(-vocab // world)
    return -(-per // 128) * 128


def validate_tp_config(world: int, n_heads: int, n_kv_heads: int) -> None:
    """reference src/engine/tensor_parallel.rs:76-101"""
    if world <= 1:
        return
    if n_heads % world:
        raise ValueError(f"num_attention_heads ({n_heads}) must be divisible by tensor_parallel_size ({world})")
    if n_kv_heads % world:
        raise ValueError(f"num_kv_heads ({n_kv_heads}) must be divisible by tensor_parallel_size ({world})")


def plan(hidden: int, n_heads: int, n_kv_heads: int, head_dim: int, ffn: int, vocab: int, rank: int, world: int) -> TpPlan:
    validate_tp_config(world, n_heads, n_kv_heads)
    h0, h1 = ops.shard_range(n_heads, rank, world)
    k0, k1 = ops.shard_range(n_kv_heads, rank, world)
    f0, f1 = ops.shard_range(ffn, rank, world, granule=256)
    # lm_head: equal-size vocabulary shards (the all-gather of the logits needs the same count on every rank):
    # rank r owns [r * vs, min((r + 1) * vs, V)), vs = ceil(V / world) rounded up to 128 rows; the last rank's shard is
    # shorter and the tail of its logits buffer stays at -inf, so the gathered [world * vs] row indexes the vocabulary directly
    vs = vocab_shard_rows(vocab, world)
    v0, v1 = min(rank * vs, vocab), min((rank + 1) * vs, vocab)
    return TpPlan(rank, world, h1 - h0, k1 - k0, (h0 * head_dim, h1 * head_dim), (k0 * head_dim, k1 * head_dim),
                  (h0 * head_dim, h1 * head_dim), (f0, f1), (f0, f1), (v0, v1))


# ------------------------------------------------------------------------------------------------
# expert parallelism: experts are partitioned num_experts / world per rank with the same
# shard_range rule; for decode every rank runs its local selected experts on the replicated hidden state and the
# partial outputs are all-reduced (single exchange).
# ------------------------------------------------------------------------------------------------
def expert_range(num_experts: int, rank: int, world: int) -> Tuple[int, int]:
    return ops.shard_range(num_experts, rank, world)


def ep_local_slots(sel, gate_w, e0: int, e1: int):
    """sel [T, top_k] global expert ids (int32 tensor), gate_w [T, top_k] -> (bank-local ids with -1 for experts
    hosted elsewhere, gate weights with those slots zeroed).  Pure tensor ops: graph-capturable, no host sync."""
    local = (sel >= e0) & (sel < e1)
    return (sel - e0).masked_fill(~local, -1), gate_w * local.to(gate_w.dtype)
