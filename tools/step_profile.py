"""Per-kernel-class timing of a few eager CFG sampler steps (library profiler: CUDA events on the launching stream)
plus the graph-replayed step time.   python tools/step_profile.py [--batch 16] [--frames 750] [--text 64] [--layers 5]
[--heads 1] [--steps 10]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ditto_tts_b200 as D  # noqa: E402
from ditto_tts_b200 import _lib  # noqa: E402
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from _opts import apply_opts  # noqa: E402

OPTS = apply_opts()   # --opt name=value -> ditto_debug_option (before the engine is created)

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=16)
ap.add_argument("--frames", type=int, default=750)
ap.add_argument("--text", type=int, default=64)
ap.add_argument("--layers", type=int, default=5)
ap.add_argument("--heads", type=int, default=1)
ap.add_argument("--hidden", type=int, default=768)
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--precision", default="bf16")
a = ap.parse_args()
dev = torch.device("cuda:0")
B, T, S, H, K = a.batch, a.frames, a.text, a.hidden, a.steps
torch.manual_seed(0)
m = D.DiTTO(hidden_dim=H, num_layers=a.layers, num_heads=a.heads, time_dim=256, text_dim=H, diffusion_steps=K,
            precision=a.precision).to(dev)
s = D.DiTTOSampler(m, guidance_scale=3.0)
g = torch.Generator().manual_seed(1)
text = torch.randn(B, S, H, generator=g).to(dev)
x0 = torch.randn(B, T, H, generator=g).to(dev)
ctx = s._context(text, True, None, T)
n = 2 * B
graph = s.step_graph(B, T, S, True, 3.0, ctx, True, dev)
graph.reset(x0, K - 1)
for _ in range(3):
    graph.replay()
graph.reset(x0, K - 1)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(K):
    graph.replay()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / K
flops = n * (a.layers * (34.0 * T * H * H + 4.0 * T * T * H + 4.0 * T * S * H) + 4.0 * T * H * H)
print(f"graph step: {ms:.3f} ms  -> {B * T / ms * 1e3:.0f} frames/s/step, {flops / ms / 1e9:.0f} algorithmic TFLOP/s "
      f"({graph.launches_per_step} launches/step)")
from ditto_tts_b200.model import _ptr, _stream  # noqa: E402
graph.reset(x0, K - 1)
_lib.profile_start()
P = min(K, 3)
for i in range(P):   # the graph's own step, un-graphed: forward + fused CFG / DDPM update with in-kernel noise
    _lib.check(_lib.load().ditto_p_sample_rng(m.engine(), _ptr(graph.x), _ptr(graph.ctx), _ptr(graph.t), _ptr(graph.rng), 1, 3.0, B, T, S,
                                              _ptr(graph.eps), _ptr(graph.x), _ptr(graph.ws), graph.ws.numel(), 1, _stream()))
prof = _lib.profile_stop()
tot = sum(v["ms"] for v in prof.values())
print(f"{'class':<24s}{'launches/step':>14s}{'ms/step':>10s}{'us/launch':>11s}{'TFLOP/s':>9s}{'GB/s':>8s}{'share':>7s}")
for k_, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"]):
    tf = v["flops"] / v["ms"] / 1e9 if v["flops"] else 0.0
    gb = v["bytes"] / v["ms"] / 1e6 if v["bytes"] else 0.0
    print(f"{k_:<24s}{v['launches'] / P:>14.0f}{v['ms'] / P:>10.4f}{v['ms'] / v['launches'] * 1e3:>11.1f}{tf:>9.0f}{gb:>8.0f}"
          f"{v['ms'] / tot:>7.1%}")
print(f"{'sum (eager, profiled)':<24s}{'':>14s}{tot / P:>10.4f}")
print("finite:", bool(torch.isfinite(graph.x).all()))
if os.environ.get("DITTO_STEP_COUNTERS"):   # library built with -DDITTO_DBG_COUNTERS=1: where the paired GEMMs' control warps wait
    import ctypes as C
    cnt = torch.zeros(64, dtype=torch.int64, device=dev)
    _lib.check(_lib.load().ditto_debug_set_counters(C.c_void_p(cnt.data_ptr())))
    _lib.check(_lib.load().ditto_p_sample_rng(m.engine(), _ptr(graph.x), _ptr(graph.ctx), _ptr(graph.t), _ptr(graph.rng), 1, 3.0, B, T, S,
                                              _ptr(graph.eps), _ptr(graph.x), _ptr(graph.ws), graph.ws.numel(), 1, _stream()))
    torch.cuda.synchronize()
    _lib.check(_lib.load().ditto_debug_set_counters(None))
    names = {0: "store_f32", 1: "store_f32_resid (fc2, proj_out)", 2: "store_bf16", 3: "generic", 4: "geglu", 5: "qkv_rope"}
    for e, c in enumerate(cnt.view(8, 8).tolist()):
        if c[2]:
            print(f"pair GEMM epilogue {names.get(e, e)}: MMA issuer waits operands {c[0] / c[2]:.1%}, accumulator {c[1] / c[2]:.1%}; "
                  f"TMA producer waits slot {c[3] / max(c[4], 1):.1%}")
