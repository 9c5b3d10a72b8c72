# dbg = flags >> 8: bit 4 = no L2 prefetch of the residual rows; dbg >> 8 = start delay of the odd clusters (x 2048 clk)
for o in 1 2; do
 for dbg in 0 16 $((8*256)) $((15*256)) $((22*256)); do
  fl=$((dbg*256))
  echo "quad=$o dbg=$dbg n=32:  $(timeout 120 python tools/attn768_bench.py --opt flash768_quad=$o --flags $fl | tail -1)"
  echo "quad=$o dbg=$dbg n=256: $(timeout 120 python tools/attn768_bench.py --opt flash768_quad=$o --flags $fl --n 256 --iters 5 | tail -1)"
 done
done
