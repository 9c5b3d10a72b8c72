"""BASELINE.json config 5: mixed-length batch (256 utterances, 2-20 s, speech-length-predictor-sized latents), 50-step CFG
sampling, bf16, sharded by cost across the visible GPUs.  Times the packed ragged path (ditto_p_sample_ragged, one CUDA
graph replay per step) and, for comparison, the padded-to-max uniform path on the same shard.

    python tools/ragged_bench.py [--utts 256] [--world 8 --rank 0]   # one rank's shard of an 8-GPU job on one GPU
    python -m torch.distributed.run --nproc-per-node N tools/ragged_bench.py  # real N-GPU job

Throughput is counted in VALID latent frames (no credit for padding); flops from SURVEY.md 8a's per-sequence model."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ditto_tts_b200 as D  # noqa: E402
from ditto_tts_b200 import parallel  # noqa: E402
from ditto_tts_b200.ragged import RaggedBatch, RaggedStepGraph  # noqa: E402
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from _opts import apply_opts  # noqa: E402

apply_opts()   # --opt name=value -> ditto_debug_option

ap = argparse.ArgumentParser()
ap.add_argument("--utts", type=int, default=256)
ap.add_argument("--world", type=int, default=0, help="simulate this many ranks (shard of --rank) on one GPU")
ap.add_argument("--rank", type=int, default=0)
ap.add_argument("--steps", type=int, default=20)
ap.add_argument("--warmup", type=int, default=3)
ap.add_argument("--padded", type=int, default=1, help="also time the padded-to-max uniform batch")
ap.add_argument("--profile", type=int, default=0, help="per-kernel-class timing of 2 eager steps")
a = ap.parse_args()

env_world = int(os.environ.get("WORLD_SIZE", "1"))
if env_world > 1:
    import torch.distributed as dist
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    world, rank = env_world, int(os.environ["RANK"])
else:
    local = 0
    world, rank = (a.world or 1), a.rank
dev = torch.device("cuda", local)
H, L, K = 768, 5, a.steps
T_all, S_all = parallel.c5_lengths(a.utts)
mine = parallel.balance_by_cost(T_all, world, S_all)[rank]
T = [T_all[i] for i in mine]
S = [S_all[i] for i in mine]
torch.manual_seed(0)
m = D.DiTTO(hidden_dim=H, num_layers=L, num_heads=1, time_dim=256, text_dim=H, diffusion_steps=K, precision="bf16").to(dev)
s = D.DiTTOSampler(m, guidance_scale=3.0)
g = torch.Generator().manual_seed(1 + rank)
texts = [torch.randn(si, H, generator=g).to(dev) for si in S]
rb = RaggedBatch(m, texts, T, guided=True)
x0 = torch.randn(rb.x_rows, H, generator=g).to(dev)
graph = RaggedStepGraph(rb, 3.0, True)


def timed(replay, reset):
    reset()
    for _ in range(a.warmup):
        replay()
    reset()
    if env_world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / K


ms = timed(graph.replay, lambda: graph.reset(x0, K - 1))
finite = bool(torch.isfinite(graph.x).all())
flops = sum(2 * parallel.utterance_cost(t, si) for t, si in zip(T, S))
res = {"workload": f"C5 shard {rank}/{world}: {len(T)} utterances, {sum(T)} valid frames, {len(rb.groups)} length groups",
       "ragged_ms_per_step": ms, "launches_per_step": graph.launches_per_step,
       "valid_frames_per_s": sum(T) / ms * 1e3, "valid_algorithmic_tflops": flops / ms / 1e9, "finite": finite}
if a.profile:
    from ditto_tts_b200 import _lib
    eps = torch.empty((rb.seq_rows, H), dtype=torch.float32, device=dev)
    xe, z = x0.clone(), torch.zeros_like(x0)
    t_seq = torch.full((rb.n_seq,), K - 1, dtype=torch.int64, device=dev)
    _lib.profile_start()
    for _ in range(2):
        rb.p_sample(xe, t_seq, z, 3.0, eps, xe)
    prof = _lib.profile_stop()
    res["per_class_ms_per_step"] = {k: round(v["ms"] / 2, 4) for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])}
    res["per_class_launches_per_step"] = {k: v["launches"] // 2 for k, v in prof.items()}
if a.padded:
    Tm, Sm, B = max(T), max(S), len(T)
    text_p = torch.zeros(B, Sm, H, device=dev)
    for i, tx in enumerate(texts):
        text_p[i, : tx.shape[0]] = tx          # NOTE: padding changes the result (the reference has no masks); timing only
    ctx = s._context(text_p, True, None, Tm)
    sg = s.step_graph(B, Tm, Sm, True, 3.0, ctx, True, dev)
    xp = torch.randn(B, Tm, H, device=dev)
    msp = timed(sg.replay, lambda: sg.reset(xp, K - 1))
    res.update({"padded_ms_per_step": msp, "padded_valid_frames_per_s": sum(T) / msp * 1e3, "ragged_speedup_vs_padded": msp / ms})
if env_world > 1:
    tmax = torch.tensor([ms], device=dev, dtype=torch.float64)
    frames = torch.tensor([float(sum(T))], device=dev, dtype=torch.float64)
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    dist.all_reduce(frames)
    res["job_valid_frames_per_s"] = float(frames) / float(tmax) * 1e3
    res["job_ms_per_step_max_over_ranks"] = float(tmax)
    lat = parallel.gather_ragged_latents(rb.unpack(graph.x), mine, T_all)
    res["gathered_utterances"] = len(lat)
    dist.destroy_process_group()
if rank == 0 or env_world == 1:
    print(json.dumps(res))
