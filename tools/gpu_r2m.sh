#!/bin/bash
# Round-2 GPU call M: the conv kernel's synchronisation skeleton (all work ablated) under pair / multicast / single-CTA modes
TAG=${1:-r02m}; OUT=gpurun_out/$TAG; mkdir -p $OUT
EGAZE_CONV_PROF=1 python egocentric-gaze-prediction_b200/csrc/build.py > $OUT/build.log 2>&1; tail -1 $OUT/build.log
for cfg in "EGAZE_CONV_PAIR=1" "EGAZE_CONV_PAIR=0" "EGAZE_CONV_CLUSTER=1"; do
for ab in 15 0; do
  echo "== $cfg ABLATE=$ab"
  env $cfg EGAZE_CONV_ABLATE=$ab PROF_ONLY="${PROF_ONLY:-64->64 @224}" timeout 300 python tools/conv_prof.py 2>&1 | cut -c1-330
done
done
