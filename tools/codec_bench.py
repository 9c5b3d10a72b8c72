"""Times the hand-off kernels either side of the loop at the C2 size (CUDA events, 20 launches after 3 warm-ups):
latents -> codes (16 x 750 frames, 1024 x 768 codebook), channel pooling, MSE.   python tools/codec_bench.py [--batch 16]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ditto_tts_b200 import codec  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=16)
ap.add_argument("--frames", type=int, default=750)
a = ap.parse_args()
dev = torch.device("cuda:0")
B, T, K, D = a.batch, a.frames, 1024, 768
torch.manual_seed(0)
vq = codec.VectorQuantizer(K, D).to(dev)
lat = torch.randn(B, T, D, device=dev) * 0.05
lat4 = torch.randn(B, 2, T, D, device=dev)


def timed(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


ms = timed(lambda: codec.latents_to_codes(vq, lat, 2))
print(f"vq_encode   {B}x{T} rows, {K}x{D} codebook: {ms * 1e3:8.1f} us  {2.0 * B * T * K * D / ms / 1e9:7.1f} fp32 TFLOP/s")
ms = timed(lambda: codec.pool_latents(lat4, 1024))
print(f"pool_latents {tuple(lat4.shape)}: {ms * 1e3:8.1f} us  {lat4.numel() * 4 * 1.5 / ms / 1e6:7.0f} GB/s")
ms = timed(lambda: codec.mse_loss(lat, lat4[:, 0]))
print(f"mse_loss    {lat.numel()} elements: {ms * 1e3:8.1f} us  {lat.numel() * 8 / ms / 1e6:7.0f} GB/s")
