"""Data-parallel sharding of an evaluation sweep: one process per GPU, utterance i -> rank i mod world, no collective on
the data path; the only exchange is one all-gather of the per-utterance (edit distance, reference length) int32 pairs
(8 bytes / utterance), after which every rank finishes with the reference's own float64 numpy expressions so the
aggregate is bit-identical to a single-process ``evaluate_batch``."""
from __future__ import annotations

from typing import List, Tuple

import numpy as np
import torch
import torch.distributed as dist


def world() -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_indices(n_total: int, rank: int, world_size: int) -> List[int]:
    """Strided assignment (balances any length skew across ranks)."""
    return list(range(rank, n_total, world_size))


def gather_counts(local_counts: torch.Tensor, n_total: int) -> np.ndarray:
    """local_counts: int32 [n_local, 2] for this rank's strided shard (device tensor under NCCL, CPU tensor under gloo).
    Returns the int32 [n_total, 2] table in original utterance order on every rank."""
    rank, ws = world()
    if ws == 1:
        out = local_counts.detach().cpu().numpy().astype(np.int32)
        assert out.shape[0] == n_total
        return out
    per_rank = (n_total + ws - 1) // ws
    padded = torch.full((per_rank, 2), -1, dtype=torch.int32, device=local_counts.device)
    padded[: local_counts.shape[0]] = local_counts
    gathered = [torch.empty_like(padded) for _ in range(ws)]
    dist.all_gather(gathered, padded)
    table = np.full((n_total, 2), -1, dtype=np.int32)
    for r in range(ws):
        idx = shard_indices(n_total, r, ws)
        table[idx] = gathered[r][: len(idx)].cpu().numpy()
    assert (table[:, 1] >= 0).all(), "an utterance was not scored by any rank"
    return table
