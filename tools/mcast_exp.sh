# does TMA multicast relieve the L2 -> SM operand feed?  1-CTA GEMM kernel (128 x 256 tiles), cluster cm x cn sharing the
# B block along m / the A block along n
for shape in "24000 6144 768" "24000 768 3072"; do
  for c in "1 1" "1 2" "2 1" "2 2" "1 4" "4 1"; do
    set -- $c
    echo -n "1cta cluster $1x$2  "; python tools/gemm_bench.py $shape --iters 20 --opt no_pair=1 --opt cluster_m=$1 --opt cluster_n=$2 | tail -1
  done
  echo -n "pair           "; python tools/gemm_bench.py $shape --iters 20 | tail -1
done
