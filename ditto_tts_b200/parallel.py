"""Utterance sharding across the GPUs of one box (one process per GPU, torch.distributed).

Utterances are independent (nothing in DiTTO.forward mixes batch elements), so the denoising loop
needs NO collective: rank r samples its shard and the final latents are gathered once (NCCL all_gather over
NVLink on B200; gloo on CPU in the tests).  Mixed-length batches are balanced by the forward cost model."""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch
import torch.distributed as dist

__all__ = ["shard_indices", "balance_by_cost", "utterance_cost", "gather_latents", "gather_ragged_latents", "c5_lengths"]


def shard_indices(n_utts: int, world: int, rank: int) -> List[int]:
    """Round-robin shard: rank r takes utterances r, r+world, ...  (equal-length batches)."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    return list(range(rank, n_utts, world))


def utterance_cost(T: int, S: int, hidden: int = 768, layers: int = 5) -> float:
    """Forward flops of one sequence: L(34 T H^2 + 4 T^2 H + 4 T S H) + 4 T H^2."""
    H = hidden
    return layers * (34.0 * T * H * H + 4.0 * T * T * H + 4.0 * T * S * H) + 4.0 * T * H * H


def balance_by_cost(lengths: Sequence[int], world: int, text_lengths: Optional[Sequence[int]] = None,
                    hidden: int = 768, layers: int = 5) -> List[List[int]]:
    """Longest-processing-time-first assignment of mixed-length utterances to ranks.  Returns, per rank, the
    utterance indices sorted by length (so that equal lengths can be batched together)."""
    S = text_lengths if text_lengths is not None else [64] * len(lengths)
    cost = [utterance_cost(int(t), int(s), hidden, layers) for t, s in zip(lengths, S)]
    order = sorted(range(len(lengths)), key=lambda i: -cost[i])
    load = [0.0] * world
    out: List[List[int]] = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: (load[k], k))
        out[r].append(i)
        load[r] += cost[i]
    for r in range(world):
        out[r].sort(key=lambda i: (lengths[i], i))
    return out


def gather_latents(local: torch.Tensor, indices: Sequence[int], n_total: int, group=None) -> torch.Tensor:
    """All-gather per-rank latents [b_r, T, H] (same T, H on every rank; b_r may differ by one) and put them
    back in utterance order.  The only collective of a sampling job."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        out = local.new_empty((n_total,) + tuple(local.shape[1:]))
        out[torch.as_tensor(list(indices), device=local.device)] = local
        return out
    world = dist.get_world_size(group)
    dev = local.device
    counts = torch.zeros(world, dtype=torch.int64, device=dev)
    counts[dist.get_rank(group)] = local.shape[0]
    dist.all_reduce(counts, group=group)
    bmax = int(counts.max())
    pad = local.new_zeros((bmax,) + tuple(local.shape[1:]))
    pad[: local.shape[0]] = local
    idx = torch.full((bmax,), -1, dtype=torch.int64, device=dev)
    idx[: local.shape[0]] = torch.as_tensor(list(indices), dtype=torch.int64, device=dev)
    all_lat = local.new_empty((world * bmax,) + tuple(local.shape[1:]))
    all_idx = torch.empty(world * bmax, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(all_lat, pad, group=group)
    dist.all_gather_into_tensor(all_idx, idx, group=group)
    keep = all_idx >= 0
    out = local.new_empty((n_total,) + tuple(local.shape[1:]))
    out[all_idx[keep]] = all_lat[keep]
    return out


def c5_lengths(n_utts: int = 256, seed: int = 0, fps: int = 75, tokens_per_second: float = 6.4):
    """BASELINE.json config 5 / SURVEY.md 8d: durations ``random.seed(0); randint(2, 20)`` s -> T_i = 75 sec frames
    (what the speech-length predictor's classes map to), S_i = round(6.4 sec) text tokens."""
    import random
    rnd = random.Random(seed)
    secs = [rnd.randint(2, 20) for _ in range(n_utts)]
    return [fps * s for s in secs], [int(round(tokens_per_second * s)) for s in secs]


def gather_ragged_latents(local: Sequence[torch.Tensor], indices: Sequence[int], lengths: Sequence[int], group=None):
    """Mixed-length counterpart of gather_latents: ``local[j]`` is the [T_i, H] latent of utterance ``indices[j]``;
    ``lengths`` holds T_i of ALL utterances (known on every rank: it is what the shards were balanced with).
    Returns the full list in utterance order on every rank.  One padded all_gather of the packed rows."""
    n_total = len(lengths)
    if len(local) != len(indices):
        raise ValueError("one latent per local index expected")
    for j, i in enumerate(indices):
        if local[j].shape[0] != lengths[i]:
            raise ValueError(f"utterance {i}: latent has {local[j].shape[0]} frames, expected {lengths[i]}")
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        out = [None] * n_total
        for j, i in enumerate(indices):
            out[i] = local[j]
        return out
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    dev, H = local[0].device, local[0].shape[1]
    # every rank's index list (padded with -1) so that the packed rows can be cut apart again
    cnt = torch.zeros(world, dtype=torch.int64, device=dev)
    cnt[rank] = len(indices)
    dist.all_reduce(cnt, group=group)
    bmax = int(cnt.max())
    idx = torch.full((bmax,), -1, dtype=torch.int64, device=dev)
    idx[: len(indices)] = torch.as_tensor(list(indices), dtype=torch.int64, device=dev)
    all_idx = torch.empty(world * bmax, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(all_idx, idx, group=group)
    all_idx = all_idx.view(world, bmax).tolist()
    rows = [sum(lengths[i] for i in r if i >= 0) for r in all_idx]
    rmax = max(rows)
    pack = local[0].new_zeros((rmax, H))
    pack[: rows[rank]] = torch.cat(list(local), dim=0)
    everything = local[0].new_empty((world * rmax, H))
    dist.all_gather_into_tensor(everything, pack, group=group)
    out = [None] * n_total
    for r in range(world):
        off = r * rmax
        for i in all_idx[r]:
            if i < 0:
                continue
            out[i] = everything[off: off + lengths[i]]
            off += lengths[i]
    return out
