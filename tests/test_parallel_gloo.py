"""Multi-process host logic on CPU: utterance sharding + the single output gather, world_size 2, gloo."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ditto_tts_b200 import parallel


def test_shard_and_balance():
    assert parallel.shard_indices(10, 4, 1) == [1, 5, 9]
    cover = sorted(i for r in range(8) for i in parallel.shard_indices(256, 8, r))
    assert cover == list(range(256))
    import random
    random.seed(0)
    lengths = [75 * random.randint(2, 20) for _ in range(256)]          # BASELINE config 5 lengths
    assert sum(lengths) == 206025
    parts = parallel.balance_by_cost(lengths, 8, [round(6.4 * t / 75) for t in lengths])
    assert sorted(i for p in parts for i in p) == list(range(256))
    loads = [sum(parallel.utterance_cost(lengths[i], 64) for i in p) for p in parts]
    assert max(loads) / (sum(loads) / 8) < 1.02                          # within 2 % of perfect balance


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n_total, T, H = 5, 3, 4
    full = torch.arange(n_total * T * H, dtype=torch.float32).reshape(n_total, T, H)
    idx = parallel.shard_indices(n_total, world, rank)                   # rank 0: 3 utterances, rank 1: 2
    local = full[idx] * 1.0
    out = parallel.gather_latents(local, idx, n_total)
    ret[rank] = bool(torch.equal(out, full))
    dist.destroy_process_group()


def test_gather_latents_world2_gloo():
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29600 + os.getpid() % 300
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    assert ret[0] and ret[1]


def test_gather_single_process():
    x = torch.randn(3, 2, 4)
    out = parallel.gather_latents(x, [2, 0, 1], 3)
    assert torch.equal(out[2], x[0]) and torch.equal(out[0], x[1])


def _ragged_worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lengths, H = [3, 7, 2, 5, 4], 4
    g = torch.Generator().manual_seed(0)
    full = [torch.randn(T, H, generator=g) for T in lengths]
    parts = parallel.balance_by_cost(lengths, world, [2] * len(lengths), hidden=H, layers=1)
    idx = parts[rank]
    out = parallel.gather_ragged_latents([full[i].clone() for i in idx], idx, lengths)
    ret[rank] = all(torch.equal(a, b) for a, b in zip(out, full))
    dist.destroy_process_group()


def test_gather_ragged_latents_world2_gloo():
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29900 + os.getpid() % 90
    mp.spawn(_ragged_worker, args=(2, port, ret), nprocs=2, join=True)
    assert ret[0] and ret[1]


def test_c5_workload_matches_survey():
    T, S = parallel.c5_lengths()
    assert len(T) == 256 and sum(T) == 206025 and min(T) == 150 and max(T) == 1500
    assert all(s == round(6.4 * t / 75) for t, s in zip(T, S))
    out = parallel.gather_ragged_latents([torch.zeros(3, 2), torch.ones(1, 2)], [1, 0], [1, 3])
    assert out[0].shape[0] == 1 and out[1].shape[0] == 3
