#!/bin/bash
# ncu --set full with source import for one kernel; exports the per-instruction SASS page (stall samples) as CSV.
# usage: tools/ncu_source.sh <tag> <name> <kernel regex> <python driver + args...>
TAG=$1; NAME=$2; K=$3; shift 3
OUT=gpurun_out; mkdir -p $OUT
timeout 300 ncu --set full --clock-control none --import-source on -k regex:$K -s 3 -c 1 -f -o /tmp/${TAG}_$NAME python "$@" > $OUT/${TAG}_src_$NAME.log 2>&1
echo "ncu rc=$?"
ncu -i /tmp/${TAG}_$NAME.ncu-rep --page source --csv --print-source sass > $OUT/${TAG}_sass_$NAME.csv 2>/dev/null
ls -la $OUT/${TAG}_sass_$NAME.csv
