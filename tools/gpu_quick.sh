#!/bin/bash
# Quick GPU-box session: parity tests, smoke, bench line, in-library per-kernel step profile.   outputs -> gpurun_out/<tag>_*
TAG=${1:-q}
O=gpurun_out
mkdir -p $O
echo "##### pytest -m gpu"
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee $O/${TAG}_pytest.log
echo "##### smoke"
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -5 | tee $O/${TAG}_smoke.log
echo "##### bench"
timeout 600 python bench.py 2> $O/${TAG}_bench.err | tee $O/${TAG}_bench.json | cut -c1-1500
tail -3 $O/${TAG}_bench.err
echo "##### step profile"
timeout 300 python tools/step_profile.py --steps 5 2>&1 | tail -22 | tee $O/${TAG}_step_profile.log
