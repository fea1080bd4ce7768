#!/bin/bash
OUT=gpurun_out/${1:-epi}; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; tail -4 $OUT/pytest_gpu.log
EGAZE_CONV_WINDOW=1 EGAZE_CONV_WINDOW_MINSB=2 timeout 600 python -m pytest tests/test_gpu_conv.py tests/test_gpu_backward.py tests/test_gpu_models.py -q -m gpu > $OUT/pytest_win1.log 2>&1; tail -4 $OUT/pytest_win1.log
for cfg in "0 3" "1 3" "1 2"; do
  set -- $cfg
  EGAZE_CONV_WINDOW=$1 EGAZE_CONV_WINDOW_MINSB=$2 timeout 300 python tools/layer_table.py > $OUT/layer_table_w$1_sb$2.txt 2>&1
  tail -1 $OUT/layer_table_w$1_sb$2.txt
done
