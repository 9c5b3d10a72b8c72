#!/usr/bin/env python
"""bench.py -- headline benchmark of the DiTTo-TTS denoiser hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

Workload (BASELINE.json configs[1], "C2"): the repo-default DiT (hidden 768, 5 layers, 1 head, time 256),
random-init weights, DDPM sampling with classifier-free guidance (w = 3, unconditional = zero text
embedding), batch 16 x 10 s utterances (T = 750 latent frames, S = 64 text tokens) PER GPU, bf16 tensor-core
path with fp32 accumulation / residual / statistics.  A "step" is one denoising step: one forward over
2B sequences (conditional + unconditional) + the fused CFG-combine/DDPM-update kernel + the step's noise draw.
K steps = one full sampling job with DIFFUSION_STEPS = K (the reference's own way of choosing the step count).

metric  = latent frames / s per denoising step = N * B * T / (time of one step), inputs resident in HBM.
e2e     = the same metric through the public API (DiTTOSampler.sample_latents) starting from PINNED HOST text
          embeddings and x_T and ending with the final latents back in host memory; copies are inside the timed
          region (the per-step noise is drawn on the device, as the reference's randn_like does).
roofline= the dominant kernel (tcgen05 GEMM with the fused GELU*sigmoid-gate epilogue) timed live with CUDA
          events on the launching stream (library profiler), algorithmic flops / measured peak.
cpu_baseline / --impl reference = the CPU restatement of the reference (oracle/, all host threads) on a bounded
          sample of the same workload (one utterance of the batch).
Parity bars (tests/, BASELINE.json): rel-L2 <= 2e-2 (bf16 path) and <= 1e-4 (fp32 path) vs the fp32 reference.
"""
from __future__ import annotations

import argparse
import atexit
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

HIDDEN, LAYERS, HEADS, TIME_DIM = 768, 5, 1, 256
FRAMES_10S, TEXT_LEN, GUIDANCE = 750, 64, 3.0


def forward_flops(T, S):
    """algorithmic flops per sequence per forward (dead self-attn out_proj excluded)."""
    H, L = HIDDEN, LAYERS
    return L * (34.0 * T * H * H + 4.0 * T * T * H + 4.0 * T * S * H) + 4.0 * T * H * H


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return dict(burst=float(d["bf16_tflops"]), sustained=float(d.get("bf16_tflops_sustained", d["bf16_tflops"])),
                    hbm=float(d["hbm_gbs"]), source="measured (MEASURED_PEAKS.json)")
    return dict(burst=1590.0, sustained=1400.0, hbm=6650.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 50 ms.  Started well before the timed region (nvidia-smi needs up
    to a second to come up on an 8-GPU box, the timed region of 50 steps lasts ~150 ms); summary() keeps the samples whose
    arrival time falls inside the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
            atexit.register(self.stop)   # never leave the sampler behind, whatever path the bench exits through
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def stop(self):
        if self.proc is not None and self.proc.poll() is None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except subprocess.TimeoutExpired:
                self.proc.kill()

    def summary(self, t0, t1):
        """Samples that arrived in [t0, t1 + one period]; when the region was shorter than the sampling period, the sample
        nearest to it (taken under the same load: warm-up steps precede and the eager profile follows the timed region)."""
        inside = [r for ts, r in self.rows if t0 <= ts <= t1 + 0.06]
        window = "timed region"
        if not inside and self.rows:
            mid = 0.5 * (t0 + t1)
            inside = [min(self.rows, key=lambda tr: abs(tr[0] - mid))[1]]
            window = "nearest sample to the timed region"
        sm, mx, reasons = [], [], set()
        for r in inside:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm),
                "window": window}


# =====================================================================================================
# reference arm / cpu baseline: the oracle port on the host cores
# =====================================================================================================
def cpu_cfg_steps(steps, warmup, max_seconds=None):
    """Times `steps` CFG denoising steps of ONE utterance (B=1, T=750, S=64, fp32) with the CPU oracle."""
    import torch
    from oracle import ditto_oracle as O  # checker / CPU baseline only
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = O.OracleConfig(HIDDEN, LAYERS, HEADS, TIME_DIM, HIDDEN, max(steps, 2))
    sd = O.make_state_dict(cfg, 0)
    x, text, _ = O.make_inputs(1, FRAMES_10S, TEXT_LEN, cfg, 1)
    betas, alphas, acp = O.sampler_tables(cfg.diffusion_steps)
    g = torch.Generator().manual_seed(2)
    done, t_total = 0, 0.0
    with torch.no_grad():
        for i in range(warmup + steps):
            t_val = cfg.diffusion_steps - 1 - (i % cfg.diffusion_steps)
            t = torch.full((1,), t_val, dtype=torch.long)
            t0 = time.perf_counter()
            eps = O.predict_noise(sd, cfg, x, text, t, GUIDANCE)
            z = torch.randn(x.shape, generator=g)
            x = O.p_sample_update(x, eps, z, t, betas, alphas, acp)
            dt = time.perf_counter() - t0
            if i >= warmup:
                done += 1
                t_total += dt
                if max_seconds is not None and t_total > max_seconds:
                    break
            if not torch.isfinite(x).all():  # random-init latents blow up after many steps; restart the state
                x, _, _ = O.make_inputs(1, FRAMES_10S, TEXT_LEN, cfg, 1)
    return done, t_total, cores


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    done, t_total, cores = cpu_cfg_steps(args.steps, args.warmup)
    ms = t_total / done * 1e3
    value = FRAMES_10S / (ms / 1e3)
    sample = (f"{done} CFG denoising steps of ONE utterance of the batch (B=1, T={FRAMES_10S}, S={TEXT_LEN}, fp32, "
              f"2 forwards + update per step), torch CPU, {cores} threads")
    line = {
        "impl": "reference", "metric": "latent frames/sec per denoising step", "value": value, "unit": "frames/s",
        "n_gpus": args.gpus, "steps": done, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, 1),
        "cpu_baseline": {"value": value, "unit": "frames/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def workload_config(args, n):
    return {"workload": f"C2: 50-step-style CFG DDPM sampling, batch {args.batch} x 10 s utterances per GPU "
                        f"(T={FRAMES_10S} frames, S={TEXT_LEN} text tokens), repo-default DiT (H=768, L=5, heads=1)",
            "batch_per_gpu": args.batch, "frames": FRAMES_10S, "text_tokens": TEXT_LEN, "guidance_scale": GUIDANCE,
            "diffusion_steps": args.steps, "sharding": f"utterances x{n} (no in-step collective)",
            "l2": "working set per step (~0.7 GB) exceeds the 126 MB L2; no explicit flush",
            "parity_bar_rel_l2": {"bf16": 2e-2, "fp32": 1e-4}}


# =====================================================================================================
# our arm
# =====================================================================================================
def run_ours(args):
    import torch
    import torch.distributed as dist
    import ditto_tts_b200 as D
    from ditto_tts_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback (use --impl reference)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    B, T, S, K, W = args.batch, FRAMES_10S, TEXT_LEN, args.steps, args.warmup
    clocks = ClockSampler(local).start()
    torch.manual_seed(0)  # random-init weights of the named architecture (default torch initialisers)
    model = D.DiTTO(hidden_dim=HIDDEN, num_layers=LAYERS, num_heads=HEADS, time_dim=TIME_DIM, text_dim=HIDDEN,
                    diffusion_steps=K, precision=args.precision).to(dev)
    sampler = D.DiTTOSampler(model, guidance_scale=GUIDANCE)
    g = torch.Generator().manual_seed(1 + rank)
    text_host = torch.randn(B, S, HIDDEN, generator=g).pin_memory()
    x_host = torch.randn(B, T, HIDDEN, generator=g).pin_memory()
    out_host = torch.empty(B, T, HIDDEN).pin_memory()

    # ---------------- device-resident loop (metric `value`) ----------------
    # the product's own stepping: one CUDA-graph replay per denoising step (noise draw + 2B-sequence forward +
    # fused CFG/DDPM update in place + step-index decrement), see ditto_tts_b200/sampler.py:StepGraph
    text = text_host.to(dev)
    x0 = x_host.to(dev)
    ctx = sampler._context(text, True, None, T)
    n = 2 * B
    graph = sampler.step_graph(B, T, S, True, GUIDANCE, ctx, True, dev)
    graph.reset(x0, K - 1)
    for i in range(W):
        graph.replay()
    graph.reset(x0, K - 1)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.perf_counter()
    e0.record()
    for i in range(K):
        graph.replay()
    e1.record()
    barrier()
    t_wall1 = time.perf_counter()
    ms_total = e0.elapsed_time(e1)
    launches = graph.launches_per_step * K      # kernels of libditto_b200 inside the K replayed graphs
    xa = graph.x
    finite = bool(torch.isfinite(xa).all())

    # eager (un-graphed) stepping, used for the per-kernel-class timing below
    t_all = torch.arange(K - 1, -1, -1, device=dev, dtype=torch.int64).unsqueeze(1).repeat(1, n).contiguous()
    eps = torch.empty((n, T, HIDDEN), dtype=torch.float32, device=dev)
    xe, z = x0.clone(), torch.empty_like(x0)

    def one_step(i):
        z.normal_()
        sampler._p_sample_raw(xe, ctx, t_all[i % K], z, True, GUIDANCE, S, eps, xe)

    # ---------------- end to end through the public API, host buffers ----------------
    def e2e_once():
        t_dev = text_host.to(dev, non_blocking=True)
        x_dev = x_host.to(dev, non_blocking=True)
        res = sampler.sample_latents(t_dev, x_init=x_dev)
        out_host.copy_(res, non_blocking=True)

    e2e_once()  # warm (allocations of the public path)
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    e2e_once()
    f1.record()
    barrier()
    ms_e2e = f0.elapsed_time(f1)

    # ---------------- per-kernel-class timing (CUDA events on the launching stream) ----------------
    _lib.profile_start()
    prof_steps = min(K, 3)
    for i in range(prof_steps):
        one_step(i)
    prof = _lib.profile_stop()
    clocks.stop()

    times = torch.tensor([ms_total, ms_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
        # NCCL is used only to gather the outputs; timed separately from the step metric
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        gathered = torch.empty((world * B, T, HIDDEN), dtype=torch.float32, device=dev)
        dist.all_gather_into_tensor(gathered, xa)
        barrier()
        g0.record()
        dist.all_gather_into_tensor(gathered, xa)
        g1.record()
        barrier()
        gather_ms = g0.elapsed_time(g1)
    else:
        gather_ms = 0.0
    ms_total, ms_e2e = float(times[0]), float(times[1])

    if rank == 0:
        peaks = load_peaks()
        ms_step = ms_total / K
        value = world * B * T / (ms_step / 1e3)
        flops_step = n * forward_flops(T, S)
        tot_prof_ms = sum(v["ms"] for v in prof.values()) or 1.0
        dom_name = max(prof, key=lambda k_: prof[k_]["ms"])
        dom = prof[dom_name]
        tc_ms = sum(v["ms"] for k_, v in prof.items() if k_.startswith("tc_gemm"))
        tc_flops = sum(v["flops"] for k_, v in prof.items() if k_.startswith("tc_gemm"))
        ach = dom["flops"] / (dom["ms"] / 1e3) / 1e12 if dom["flops"] else 0.0
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            with open(tp) as f:
                traffic = json.load(f).get(dom_name)
        roofline = {"kernel": dom_name, "bound": "tensor", "achieved": ach, "peak": peaks["sustained"], "unit": "TFLOP/s",
                    "frac": ach / peaks["sustained"], "traffic": traffic, "peak_source": peaks["source"] + ", sustained",
                    "launch_us": dom["ms"] / dom["launches"] * 1e3, "flops_per_launch": dom["flops"] / dom["launches"],
                    "share_of_step": dom["ms"] / tot_prof_ms,
                    "tc_gemm_family": {"achieved": tc_flops / (tc_ms / 1e3) / 1e12 if tc_ms else 0.0,
                                       "share_of_step": tc_ms / tot_prof_ms},
                    "whole_step": {"algorithmic_tflops": flops_step / (ms_step / 1e3) / 1e12,
                                   "frac_of_peak": flops_step / (ms_step / 1e3) / 1e12 / peaks["sustained"]},
                    "per_class_ms_per_step": {k_: round(v["ms"] / prof_steps, 4) for k_, v in sorted(prof.items())}}
        done, t_cpu, cores = cpu_cfg_steps(steps=12, warmup=1, max_seconds=12.0)
        cpu_val = FRAMES_10S / (t_cpu / done)
        e2e_val = world * B * T * K / (ms_e2e / 1e3)
        line = {
            "metric": "latent frames/sec per denoising step", "value": value, "unit": "frames/s", "n_gpus": world,
            "steps": K, "warmup": W, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
            "config": workload_config(args, world),
            "rtf": {"batch_latency_s": ms_total / 1e3, "rtf_10s_utterance": ms_total / 1e3 / 10.0,
                    "rtf_per_audio_second": ms_total / 1e3 / (world * B * 10.0)},
            "roofline": roofline,
            "cpu_baseline": {"value": cpu_val, "unit": "frames/s", "cores": cores, "kind": "port",
                             "sample": f"{done} CFG denoising steps of one utterance (B=1, T={T}, S={S}, fp32) with the "
                                       f"CPU oracle port of the reference, {cores} torch threads"},
            "e2e": {"value": e2e_val, "unit": "frames/s",
                    "h2d_bytes_per_step": (text_host.numel() + x_host.numel()) * 4 / K,
                    "d2h_bytes_per_step": out_host.numel() * 4 / K, "ms_total": ms_e2e,
                    "api": "DiTTOSampler.sample_latents(text_emb, x_init) from pinned host tensors, final latents to host"},
            "gpu_launches": int(launches), "launches_per_step": int(graph.launches_per_step), "stepping": "cuda-graph replay per step", "clocks": clocks.summary(t_wall0, t_wall1), "outputs_finite": finite,
            "output_gather_ms": gather_ms,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--batch", type=int, default=16, help="utterances per GPU")
    ap.add_argument("--precision", choices=["bf16", "fp32"], default="bf16")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        world = int(os.environ.get("WORLD_SIZE", "1"))
        if args.gpus > 1 and world == 1:
            # convenience: re-launch under torchrun
            cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
                   "--master-addr", "127.0.0.1", "--master-port", str(29500 + os.getpid() % 1000), __file__,
                   "--gpus", str(args.gpus), "--steps", str(args.steps), "--warmup", str(args.warmup),
                   "--batch", str(args.batch), "--precision", args.precision]
            raise SystemExit(subprocess.call(cmd))
        run_ours(args)


if __name__ == "__main__":
    main()
