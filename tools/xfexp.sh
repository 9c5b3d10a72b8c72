#!/bin/bash
# cross_fused_kernel: L2 prefetch distance of the residual stream (xf_prefetch: 0 = off (default), n = n half-chunk jobs ahead)
O=gpurun_out; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "fused_cross or full_size_forward" 2>&1 | tail -2
for b in 16 128; do
  for pf in 0 6 3 12; do
    echo "batch=$b xf_prefetch=$pf: $(timeout 200 python tools/step_profile.py --batch $b --steps 20 --opt xf_prefetch=$pf 2>&1 | grep 'graph step\|cross_fused' | tr '\n' ' ' | tr -s ' ')"
  done
done 2>&1 | tee $O/xfexp.txt
