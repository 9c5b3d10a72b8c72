#!/bin/bash
# Round-2 GPU-box session: parity tests, smoke, bench lines (default = C4, plus C3 / C5), reference arm, per-class step profile.
# Usage (under gpurun): bash tools/gpu_round2.sh [tag] [quick]      outputs -> gpurun_out/<tag>_*
TAG=${1:-r2}
QUICK=${2:-}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/${TAG}_gpu.txt
echo "##### pytest -m gpu"
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -25 | tee $O/${TAG}_pytest.log
echo "##### smoke"
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -5 | tee $O/${TAG}_smoke.log
echo "##### bench (default: C4 on one GPU + C2 beside it)"
timeout 900 python bench.py 2> $O/${TAG}_bench.err | tee $O/${TAG}_bench.json | cut -c1-2500
tail -5 $O/${TAG}_bench.err
echo "##### step profile C2"
timeout 300 python tools/step_profile.py --steps 10 2>&1 | tail -24 | tee $O/${TAG}_step_profile.log
if [ "$QUICK" != "quick" ]; then
echo "##### bench c3 / c5"
timeout 600 python bench.py --workload c3 2>> $O/${TAG}_bench.err | tee $O/${TAG}_bench_c3.json | cut -c1-600
timeout 900 python bench.py --workload c5 2>> $O/${TAG}_bench.err | tee $O/${TAG}_bench_c5.json | cut -c1-600
echo "##### reference arm"
timeout 600 python bench.py --impl reference --steps 10 --warmup 1 | tee $O/${TAG}_bench_ref.json | cut -c1-900
fi
ls -la $O | tail -20
