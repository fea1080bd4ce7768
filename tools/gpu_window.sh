#!/bin/bash
# Probe the swizzle-shift addressing and test the conv "window" mode (correctness + per-layer timing).
OUT=gpurun_out/${1:-win}; mkdir -p $OUT
timeout 120 ./tools/probe_swizzle_shift > $OUT/probe.txt 2>&1; echo "probe exit $?" >> $OUT/probe.txt
grep -c " ok " $OUT/probe.txt; grep -c MISMATCH $OUT/probe.txt; grep "img=0" $OUT/probe.txt | head -40
for mode in 1 2; do
  EGAZE_CONV_WINDOW=$mode EGAZE_CONV_WINDOW_MINSB=2 timeout 600 python -m pytest tests/test_gpu_conv.py tests/test_gpu_backward.py -q -m gpu -k "conv3x3 or dgrad" > $OUT/pytest_win$mode.log 2>&1
  tail -4 $OUT/pytest_win$mode.log
done
for cfg in "0 3" "1 3" "1 2" "2 3" "2 2"; do
  set -- $cfg
  EGAZE_CONV_WINDOW=$1 EGAZE_CONV_WINDOW_MINSB=$2 timeout 300 python tools/layer_table.py > $OUT/layer_table_w$1_sb$2.txt 2>&1
  tail -1 $OUT/layer_table_w$1_sb$2.txt
done
