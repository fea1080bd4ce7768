#!/bin/bash
# Round-2 GPU call O: compact MMA tap loop -- conv tests, per-layer table (lean vs round-1 loop), skeleton ablation
TAG=${1:-r02o}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_conv.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -2
for pb in ${PROBES:-0 1}; do
  EGAZE_CONV_PROBE=$pb timeout 300 python tools/layer_table.py > $OUT/layer_table_probe$pb.txt 2>&1
  echo "PROBE=$pb: $(tail -1 $OUT/layer_table_probe$pb.txt)"
done
paste <(awk '{print $1,$3,$4,$5,$8}' $OUT/layer_table_probe0.txt) <(awk '{print $8}' $OUT/layer_table_probe1.txt) | awk '$1!="timed"{k=$1" "$2" "$3" "$4; n[k]++; a[k]+=$5; b[k]+=$6} END{for(k in n) printf "%-40s x%2d lean %.3f old %.3f\n", k, n[k], a[k]/n[k], b[k]/n[k]}' | sort
EGAZE_CONV_PROF=1 python egocentric-gaze-prediction_b200/csrc/build.py > $OUT/build.log 2>&1; tail -1 $OUT/build.log
for ab in 0 15; do
  echo "== ABLATE=$ab"
  EGAZE_CONV_ABLATE=$ab PROF_ONLY="${PROF_ONLY:-64 @224}" timeout 300 python tools/conv_prof.py 2>&1 | cut -c1-330
done
