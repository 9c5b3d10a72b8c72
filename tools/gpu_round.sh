#!/bin/bash
# One GPU-box session: GPU parity tests, smoke, bench (ours + reference arm), ncu launch list, ncu --set full of the top kernels.
# Usage (under gpurun): bash tools/gpu_round.sh [tag]      outputs -> gpurun_out/<tag>_*
TAG=${1:-r1}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/${TAG}_gpu.txt
echo "##### pytest -m gpu"
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee $O/${TAG}_pytest.log
echo "##### smoke"
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -5 | tee $O/${TAG}_smoke.log
echo "##### bench"
timeout 600 python bench.py 2> $O/${TAG}_bench.err | tee $O/${TAG}_bench.json
tail -5 $O/${TAG}_bench.err
if [ "${SKIP_NCU:-0}" != "1" ]; then
echo "##### ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/${TAG}_launches.csv \
    python bench.py --steps 3 --warmup 3 > $O/${TAG}_ncu_bench.log 2>&1
tail -2 $O/${TAG}_ncu_bench.log
echo "##### ncu full"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'tc_gemm|softmax|adaln|layernorm|cfg_ddpm' -s 120 -c 40 \
    -o $O/${TAG}_prof -f python bench.py --steps 3 --warmup 3 > $O/${TAG}_ncu_full.log 2>&1
tail -2 $O/${TAG}_ncu_full.log
fi
echo "##### reference arm"
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 | tee $O/${TAG}_bench_ref.json
ls -la $O
