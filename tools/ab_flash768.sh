timeout 300 python -m pytest tests/test_gpu_flash768.py -q 2>&1 | tail -3
timeout 120 python tools/attn768_bench.py --n 256 --T 750 --iters 5
timeout 120 python tools/attn768_bench.py
echo "#### C2 step, flash768 on"; python tools/step_profile.py --steps 20 2>&1 | grep "graph step\|flash768\|self_\|layernorm"
echo "#### C2 step, flash768 off"; python tools/step_profile.py --steps 20 --opt no_flash768=1 2>&1 | grep "graph step\|flash768\|self_\|layernorm"
echo "#### C4-size step (B=128), flash768 on"; python tools/step_profile.py --steps 20 --batch 128 2>&1 | grep "graph step\|flash768\|self_\|layernorm"
echo "#### C4-size step (B=128), flash768 off"; python tools/step_profile.py --steps 20 --batch 128 --opt no_flash768=1 2>&1 | grep "graph step\|flash768\|self_\|layernorm"
