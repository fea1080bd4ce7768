#!/bin/bash
# Round-2 GPU call N: ncu --set full of the 64-channel 224^2 layers: warp stall reasons (is it instruction fetch?)
TAG=${1:-r02n}; OUT=gpurun_out/$TAG; mkdir -p $OUT
REPS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:'conv3x3_tc' -c 7 -f -o $OUT/prof64 python tools/ncu_conv.py > $OUT/ncu.log 2>&1; tail -3 $OUT/ncu.log
ncu -i $OUT/prof64.ncu-rep --page raw --csv > $OUT/prof64_raw.csv 2>/dev/null
ncu -i $OUT/prof64.ncu-rep --page details --csv > $OUT/prof64_details.csv 2>/dev/null
ncu -i $OUT/prof64.ncu-rep --page source --csv --launch-skip 0 --launch-count 1 > $OUT/prof64_source0.csv 2>/dev/null
ncu -i $OUT/prof64.ncu-rep --page source --csv --launch-skip 5 --launch-count 1 > $OUT/prof64_source5.csv 2>/dev/null
gzip -f $OUT/prof64_source*.csv
rm -f $OUT/prof64.ncu-rep
ls -la $OUT
