#!/bin/bash
# ncu --set full capture of selected kernels of one profiled step; raw-metric CSV and (optionally) the report go to gpurun_out/.
# Usage (under gpurun): bash tools/ncu_capture.sh <tag> <kernel-regex> <skip> <count> [cmd...]
TAG=$1; RE=$2; SKIP=$3; CNT=$4; shift 4
CMD=${@:-python tools/step_profile.py --steps 3}
O=gpurun_out; mkdir -p $O
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"$RE" -s $SKIP -c $CNT -f -o /tmp/${TAG} $CMD > $O/${TAG}_ncu.log 2>&1
tail -3 $O/${TAG}_ncu.log
ncu -i /tmp/${TAG}.ncu-rep --page raw --csv > $O/${TAG}_raw.csv 2>/dev/null
SZ=$(stat -c %s /tmp/${TAG}.ncu-rep)
echo "report size $SZ"
if [ "$SZ" -lt 45000000 ]; then cp /tmp/${TAG}.ncu-rep $O/; fi
