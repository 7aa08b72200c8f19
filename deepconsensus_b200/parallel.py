"""Work split for multi-GPU runs: windows are independent units, so a ZMW batch is sharded across
ranks with no data-path collective.

Shards are ZMW-granular (all windows of a molecule stay on one rank so `stitch_to_fastq`, which
runs per ZMW after the model -- quick_inference.py:721-736 -- needs no cross-rank merge) and assigned
round-robin in order of first appearance.  Only counters are ever reduced across ranks.
"""
from __future__ import annotations

from typing import Dict, List, Sequence


def shard_by_zmw(molecule_names: Sequence[str], world_size: int) -> List[List[int]]:
  """Window indices per rank; molecules are dealt round-robin in order of first appearance."""
  if world_size <= 0:
    raise ValueError("world_size must be positive")
  owner: Dict[str, int] = {}
  shards: List[List[int]] = [[] for _ in range(world_size)]
  for i, name in enumerate(molecule_names):
    if name not in owner:
      owner[name] = len(owner) % world_size
    shards[owner[name]].append(i)
  return shards


def shard_range(n_items: int, rank: int, world_size: int) -> range:
  """Contiguous, balanced split of n independent items (used for synthetic window batches)."""
  base, rem = divmod(n_items, world_size)
  start = rank * base + min(rank, rem)
  return range(start, start + base + (1 if rank < rem else 0))


def reduce_counters(counters: Dict[str, int], group=None) -> Dict[str, int]:
  """SUM-all-reduce outcome counters across ranks (the only collective on the path)."""
  import torch
  import torch.distributed as dist
  if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
    return dict(counters)
  keys = sorted(counters)
  dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
  t = torch.tensor([counters[k] for k in keys], dtype=torch.int64, device=dev)
  dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
  return {k: int(v) for k, v in zip(keys, t.tolist())}
