#!/bin/bash
# Round-2 GPU call R: conv tests + per-layer table of the current build
TAG=${1:-r02r}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_conv.py -m gpu -q -x -p no:cacheprovider > $OUT/test_conv.log 2>&1; rc=$?; tail -3 $OUT/test_conv.log | cut -c1-200; grep -E "^E  " $OUT/test_conv.log | head -10 | cut -c1-220
if [ $rc -ne 0 ]; then exit 0; fi
for rep in 1 2; do
EGAZE_CONV_DIRECT=0 timeout 300 python tools/layer_table.py > $OUT/layer_table_$rep.txt 2>&1
echo "run $rep: $(tail -1 $OUT/layer_table_$rep.txt)"
done
awk '{print $1,$3,$4,$5,$7,$8}' $OUT/layer_table_1.txt | awk '$1!="timed"{k=$1" "$2" "$3" "$4" "$5; n[k]++; a[k]+=$6} END{for(k in n) printf "%-48s x%2d %.3f\n", k, n[k], a[k]/n[k]}' | sort | grep conv
