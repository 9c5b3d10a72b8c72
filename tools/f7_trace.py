"""Timeline of one cluster of flash_attn768_kernel (developer tool; needs a library built with the trace compiled in:
    DITTO_NVCC_EXTRA=-DDITTO_F7_TRACE=1 python -m ditto_tts_b200.build --force
    python tools/f7_trace.py [--n 32] [--T 750] [--items 2]
Prints, for cluster 0, the events of the MMA issuer, the first softmax warp and the TMA producer of both CTAs, in SM clocks
since the cluster barrier (the two SMs' clocks are aligned at that barrier)."""
import argparse
import ctypes as C
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from ditto_tts_b200 import _lib  # noqa: E402
from _opts import apply_opts  # noqa: E402

apply_opts()
ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=32)
ap.add_argument("--T", type=int, default=750)
ap.add_argument("--limit", type=int, default=200)
a = ap.parse_args()
dev = torch.device("cuda:0")
H = 768
P = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
g = torch.Generator(device=dev).manual_seed(0)
qkv = torch.randn(a.n * a.T, 3 * H, device=dev, generator=g).bfloat16()
h = torch.randn(a.n * a.T, H, device=dev, generator=g)
u = torch.empty(a.n * a.T, H, dtype=torch.bfloat16, device=dev)
gamma, beta = torch.ones(H, device=dev), torch.zeros(H, device=dev)
lib = _lib.load()
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
ROLES, MAX, CTAS = 3, 512, 4
buf = torch.zeros(CTAS * ROLES * MAX, dtype=torch.int64, device=dev)


def run():
    _lib.check(lib.ditto_attn_self768(P(qkv), 3 * H, a.n, a.T, 1.0 / math.sqrt(H), P(h), P(gamma), P(beta), P(u), 0, st))


for _ in range(2):
    run()
torch.cuda.synchronize()
_lib.check(lib.ditto_debug_set_counters(P(buf)))
run()
torch.cuda.synchronize()
_lib.check(lib.ditto_debug_set_counters(None))
KIND = {8: "PV.stage0_issued", 9: "PV.stage1_issued", 25: "sm.rescale_o", 26: "sm.res_loads_begin", 27: "sm.res_loads_issued", 1: "S.begin", 2: "S.issued", 3: "PV.ready", 4: "PV.issued", 5: "item.issued", 6: "S.feedwait", 7: "PV.feedwait",
        10: "sm.scores_in_regs", 11: "sm.hdr_recv", 12: "sm.decision_sent", 13: "sm.p_free", 14: "sm.P_written", 15: "sm.l_recv",
        16: "sm.o_full", 17: "sm.sweep1_done", 18: "sm.stat_recv", 19: "sm.item_done", 20: "sm.tile_begin", 21: "sm.sweep1_ld", 22: "sm.sweep1_st", 23: "sm.sweep2_ld", 24: "sm.prefetched",
        30: "tma.S_loads_issued", 31: "tma.V_loads_issued"}
ROLE = ["mma", "softmax", "tma"]
ev = []
b = buf.cpu().tolist()
for cta in range(CTAS):
    for role in range(ROLES):
        base = (cta * ROLES + role) * MAX
        for i in range(MAX):
            w = b[base + i] & 0xFFFFFFFFFFFFFFFF
            if w == 0:
                break
            tag, t = w >> 40, w & 0xFFFFFFFFFF
            ev.append((t if (tag >> 8) not in (6, 7) else -1, cta, ROLE[role], KIND.get(tag >> 8, str(tag >> 8)), tag & 255, t, i))
# durations (feed waits) are printed right after the event recorded before them
ordered = sorted([e for e in ev if e[0] >= 0])
dur = {(e[1], e[2], e[6]): e for e in ev if e[0] < 0}
print(f"{len(ev)} events; times in SM clocks since the cluster barrier")
shown = 0
for t, cta, role, kind, j, _, i in ordered:
    extra = ""
    for k in (1, 2):
        d = dur.get((cta, role, i + k))
        if d is not None and d[6] == i + k and (k == 1 or dur.get((cta, role, i + 1)) is not None or True):
            if k == 1:
                extra = f"   [{d[3]} {d[5]} clk]"
            break
    print(f"{t:>9d}  cta{cta} {role:<8s} {kind:<20s} j={j}{extra}")
    shown += 1
    if shown >= a.limit * 4:
        break
