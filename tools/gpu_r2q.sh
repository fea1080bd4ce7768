#!/bin/bash
# Round-2 GPU call Q: single-lane lean MMA issuer x direct / staged epilogue -- conv tests, per-layer table, skeleton ablation
TAG=${1:-r02q}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_conv.py -m gpu -q -x -p no:cacheprovider > $OUT/test_conv.log 2>&1; rc=$?; tail -3 $OUT/test_conv.log | cut -c1-200; grep -E "^E  " $OUT/test_conv.log | head -10 | cut -c1-220
if [ $rc -ne 0 ]; then exit 0; fi
for d in 1 0; do
  EGAZE_CONV_DIRECT=$d timeout 300 python tools/layer_table.py > $OUT/layer_table_direct$d.txt 2>&1
  echo "DIRECT=$d: $(tail -1 $OUT/layer_table_direct$d.txt)"
done
paste <(awk '{print $1,$3,$4,$5,$8}' $OUT/layer_table_direct1.txt) <(awk '{print $8}' $OUT/layer_table_direct0.txt) | awk '$1!="timed"{k=$1" "$2" "$3" "$4; n[k]++; a[k]+=$5; b[k]+=$6} END{for(k in n) printf "%-40s x%2d direct %.3f staged %.3f\n", k, n[k], a[k]/n[k], b[k]/n[k]}' | sort | grep conv
EGAZE_CONV_PROF=1 python egocentric-gaze-prediction_b200/csrc/build.py > $OUT/build.log 2>&1; tail -1 $OUT/build.log
for d in 1 0; do for ab in 0 15; do
  echo "== DIRECT=$d ABLATE=$ab"
  EGAZE_CONV_DIRECT=$d EGAZE_CONV_ABLATE=$ab PROF_ONLY="${PROF_ONLY:-64 @224}" timeout 300 python tools/conv_prof.py 2>&1 | grep -v "epilogue (thread" | cut -c1-330
done; done
