# A/B of the two self-attention kernels (flash768_quad = 1: two-CTA, 2: four-CTA) at the C2 and the C4 problem size
for o in 1 2; do
  echo "quad=$o n=32:  $(timeout 120 python tools/attn768_bench.py --opt flash768_quad=$o | tail -1)"
  echo "quad=$o n=256: $(timeout 120 python tools/attn768_bench.py --opt flash768_quad=$o --n 256 --iters 5 | tail -1)"
done
timeout 300 python -m pytest tests/test_gpu_flash768.py -x -q 2>&1 | tail -2
