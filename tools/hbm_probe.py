"""HBM directionality probe (torch library kernels, 2 GiB buffers, CUDA events): pure read, pure write, copy.  Explains why the
write-heavy kernels of the step (AdaLN: 0.6 GB read, 2.1 GB written at C4) sit far below the copy figure of MEASURED_PEAKS.json."""
import torch

dev = torch.device("cuda:0")
n = 1 << 29   # 2 GiB of fp32
a = torch.randn(n, device=dev)
b = torch.empty_like(a)


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


ms = timed(lambda: a.sum())
print(f"pure read  (sum of 2 GiB):        {ms:7.3f} ms  {a.numel() * 4 / ms / 1e6:7.0f} GB/s")
ms = timed(lambda: b.fill_(1.0))
print(f"pure write (fill of 2 GiB):       {ms:7.3f} ms  {b.numel() * 4 / ms / 1e6:7.0f} GB/s")
ms = timed(lambda: b.copy_(a))
print(f"copy       (2 GiB -> 2 GiB):      {ms:7.3f} ms  {2 * b.numel() * 4 / ms / 1e6:7.0f} GB/s (read + write)")
ms = timed(lambda: torch.add(a, 1.0, out=b))
print(f"read + write (add scalar, out=):  {ms:7.3f} ms  {2 * b.numel() * 4 / ms / 1e6:7.0f} GB/s (read + write)")
c = torch.empty(n // 2, device=dev)
ms = timed(lambda: torch.add(a[: n // 2], a[n // 2:], out=c))
print(f"2 reads : 1 write (add, out=):    {ms:7.3f} ms  {3 * c.numel() * 4 / ms / 1e6:7.0f} GB/s")
