#!/bin/bash
# Round-2 GPU call L: ablation of the conv kernel's roles on the 64-channel 224^2 layers (a -DEGAZE_CONV_PROF build made on the box):
# which of MMA issue / store loop / TMEM->smem / activation loads bounds the item time
TAG=${1:-r02l}; OUT=gpurun_out/$TAG; mkdir -p $OUT
for spin in ${SPINS:-0 1}; do
if [ $spin = 1 ]; then export EGAZE_MBAR_SPIN=1; else unset EGAZE_MBAR_SPIN; fi
EGAZE_CONV_PROF=1 python egocentric-gaze-prediction_b200/csrc/build.py > $OUT/build.log 2>&1; tail -1 $OUT/build.log
for ab in ${ABLATES:-0 15}; do
  echo "== SPIN=$spin ABLATE=$ab"
  EGAZE_CONV_ABLATE=$ab PROF_ONLY="${PROF_ONLY:-@224}" timeout 300 python tools/conv_prof.py 2>&1 | grep -v "epilogue (thread" | cut -c1-75
done
done
