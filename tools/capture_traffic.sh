#!/bin/bash
# DRAM traffic of the dominant kernel (GLU GEMM) at the bench's problem size (C4: 256 utterances on one GPU), for
# bench.py's `roofline.traffic`: one `ncu --set full` capture of a single launch inside tools/step_profile.py --batch 256.
# Usage (under gpurun): bash tools/capture_traffic.sh <tag>      then, here: python tools/make_traffic.py gpurun_out/<tag>_glu_raw.csv
TAG=${1:-r2}
O=gpurun_out; mkdir -p $O
timeout 1500 ncu --set full --clock-control none --kernel-name-base demangled -k regex:'tc_gemm_pair_kernel<\(int\)4>' -s 12 -c 1 -f \
    -o /tmp/${TAG}_glu python tools/step_profile.py --batch 256 --steps 3 > $O/${TAG}_glu_ncu.log 2>&1
tail -3 $O/${TAG}_glu_ncu.log
ncu -i /tmp/${TAG}_glu.ncu-rep --page raw --csv > $O/${TAG}_glu_raw.csv 2>/dev/null
ls -la /tmp/${TAG}_glu.ncu-rep $O/${TAG}_glu_raw.csv
