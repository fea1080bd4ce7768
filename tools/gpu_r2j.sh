#!/bin/bash
# Round-2 GPU call J: per-role / per-epilogue-phase cycle accounting of the conv kernel (a -DEGAZE_CONV_PROF build made on the box)
TAG=${1:-r02j}; OUT=gpurun_out/$TAG; mkdir -p $OUT
EGAZE_CONV_PROF=1 python egocentric-gaze-prediction_b200/csrc/build.py > $OUT/build.log 2>&1; tail -2 $OUT/build.log
for pb in ${PROBES:-1 2}; do
EGAZE_CONV_PROBE=$pb PROF_ONLY="$PROF_ONLY" timeout 600 python tools/conv_prof.py > $OUT/conv_prof_probe$pb.txt 2>&1; echo "== PROBE=$pb"; cat $OUT/conv_prof_probe$pb.txt | cut -c1-360
done
