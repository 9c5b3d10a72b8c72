#!/bin/bash
# mainloop experiments: big square GEMM and the hot shapes, 1-CTA vs pair, ring depth sweep
for shape in "8192 8192 8192" "24000 2304 768" "24000 768 3072"; do
  for st in 2 3 4 5 6; do echo -n "pair stages=$st  "; python tools/gemm_bench.py $shape --iters 20 --opt stages_pair=$st | tail -1; done
  for st in 2 3 4; do echo -n "1cta stages=$st  "; python tools/gemm_bench.py $shape --iters 20 --opt no_pair=1 --opt stages_1cta=$st | tail -1; done
done
python - <<'PY'
import torch
for sh in ((8192,8192,8192),(24000,2304,768),(24000,768,3072)):
    M,N,K=sh
    a=torch.randn(M,K,device='cuda').bfloat16(); b=torch.randn(N,K,device='cuda').bfloat16()
    for _ in range(3): c=a@b.T
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): c=a@b.T
    e1.record(); torch.cuda.synchronize()
    ms=e0.elapsed_time(e1)/20
    print(f"cuBLAS (torch.matmul) M{M} N{N} K{K}: {ms*1e3:.1f} us {2.0*M*N*K/ms/1e9:.1f} TFLOP/s")
PY
