#!/bin/bash
OUT=gpurun_out/${1:-issue}; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; tail -3 $OUT/pytest_gpu.log
run() { tag=$1; shift; env "$@" timeout 300 python tools/conv_prof.py > $OUT/conv_prof_$tag.txt 2>&1; echo "== $tag"; cut -c1-100 $OUT/conv_prof_$tag.txt; }
run base A=1
run win EGAZE_CONV_WINDOW=1 EGAZE_CONV_WINDOW_MINSB=2
for cfg in "0 3 1" "1 3 1" "1 2 1" "0 3 0"; do
  set -- $cfg
  EGAZE_CONV_WINDOW=$1 EGAZE_CONV_WINDOW_MINSB=$2 EGAZE_WGRAD_STACKED=$3 timeout 300 python tools/layer_table.py > $OUT/layer_table_w$1_sb$2_st$3.txt 2>&1
  tail -1 $OUT/layer_table_w$1_sb$2_st$3.txt
done
