"""Isolated launches of the tcgen05 GEMM for ncu captures / timing.
    python tools/gemm_bench.py [M N K] [--iters n] [--f32out] [--resid]"""
import ctypes as C
import sys

import torch

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from ditto_tts_b200 import _lib  # noqa: E402
from _opts import apply_opts  # noqa: E402

apply_opts()   # --opt name=value -> ditto_debug_option

args = [a for a in sys.argv[1:] if not a.startswith("--")]
M, N, K = (int(a) for a in args[:3]) if len(args) >= 3 else (24000, 2304, 768)
iters = int(sys.argv[sys.argv.index("--iters") + 1]) if "--iters" in sys.argv else 10
f32out = "--f32out" in sys.argv
resid = "--resid" in sys.argv
dev = torch.device("cuda:0")
lib = _lib.load()
P = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
A = torch.randn(M, K, device=dev).bfloat16()
W = torch.randn(N, K, device=dev).bfloat16()
Cd = torch.empty(M, N, dtype=torch.float32 if f32out else torch.bfloat16, device=dev)
bias = torch.randn(N, device=dev)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)


def run():
    _lib.check(lib.ditto_gemm_bf16(P(A), K, P(W), K, P(Cd), N, 0 if f32out else 1, P(bias), P(Cd) if resid else None, N, 1.0,
                                   M, N, K, st))


for _ in range(3):
    run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters):
    run()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / iters
print(f"gemm_bf16 M{M} N{N} K{K} f32out={f32out} resid={resid}: {ms * 1e3:.1f} us  {2.0 * M * N * K / ms / 1e9:.1f} TFLOP/s")
if "--counters" in sys.argv:  # where the paired kernel's MMA issuer / TMA producer spend their cycles
    cnt = torch.zeros(64, dtype=torch.int64, device=dev)   # 8 counters per epilogue kind of the paired kernel
    _lib.check(lib.ditto_debug_set_counters(P(cnt)))
    run()
    torch.cuda.synchronize()
    _lib.check(lib.ditto_debug_set_counters(None))
    c = cnt.view(8, 8).sum(0).tolist()
    print(f"  mma issuer: wait operands {c[0] / max(c[2], 1):.1%}, wait accumulator {c[1] / max(c[2], 1):.1%} of {c[2]} cycles; "
          f"tma producer: wait slot {c[3] / max(c[4], 1):.1%} of {c[4]} cycles")
