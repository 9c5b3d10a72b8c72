#!/bin/bash
# ncu --set full of ONE kernel of a tool run: bash tools/ncu_one.sh <tag> <kernel regex> <skip> <python args...>
TAG=$1; KRE=$2; SKIP=$3; shift 3
O=gpurun_out; mkdir -p $O
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$KRE" -s $SKIP -c 1 -o /tmp/${TAG} -f python "$@" > $O/${TAG}_ncu.log 2>&1
tail -3 $O/${TAG}_ncu.log
ncu -i /tmp/${TAG}.ncu-rep --page raw --csv > $O/${TAG}_raw.csv 2>/dev/null
ncu -i /tmp/${TAG}.ncu-rep --page source --csv > $O/${TAG}_source.csv 2>/dev/null
ls -la /tmp/${TAG}.ncu-rep $O/${TAG}_raw.csv $O/${TAG}_source.csv
