#!/bin/bash
# Round-2 GPU call P: direct (warp-private) epilogue -- conv tests, then the suite's conv-dependent files, per-layer table A/B
TAG=${1:-r02p}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_conv.py -m gpu -q -x -p no:cacheprovider > $OUT/test_conv.log 2>&1; rc=$?; tail -3 $OUT/test_conv.log | cut -c1-200; grep -E "^E  " $OUT/test_conv.log | head -10 | cut -c1-220
if [ $rc -ne 0 ]; then
  timeout 600 python -m pytest tests/test_gpu_conv.py -m gpu -q -p no:cacheprovider 2>&1 | grep -E "passed|failed|FAILED" | head -60 | cut -c1-200
  exit 0
fi
for t in backward models golden graph; do
  timeout 900 python -m pytest tests/test_gpu_$t.py -m gpu -q -x -p no:cacheprovider > $OUT/test_gpu_$t.log 2>&1; echo "test_gpu_$t exit $?"
  tail -2 $OUT/test_gpu_$t.log | cut -c1-200; grep -E "^E  " $OUT/test_gpu_$t.log | head -8 | cut -c1-220
done
for d in 1 0; do
  EGAZE_CONV_DIRECT=$d timeout 300 python tools/layer_table.py > $OUT/layer_table_direct$d.txt 2>&1
  echo "DIRECT=$d: $(tail -1 $OUT/layer_table_direct$d.txt)"
done
paste <(awk '{print $1,$3,$4,$5,$8}' $OUT/layer_table_direct1.txt) <(awk '{print $8}' $OUT/layer_table_direct0.txt) | awk '$1!="timed"{k=$1" "$2" "$3" "$4; n[k]++; a[k]+=$5; b[k]+=$6} END{for(k in n) printf "%-40s x%2d direct %.3f staged %.3f\n", k, n[k], a[k]/n[k], b[k]/n[k]}' | sort
