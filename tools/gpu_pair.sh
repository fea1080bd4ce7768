#!/bin/bash
OUT=gpurun_out/${1:-pair}; mkdir -p $OUT
EGAZE_CONV_PAIR=1 timeout 600 python -m pytest tests/test_gpu_conv.py -m gpu -x -q > $OUT/pytest_conv.log 2>&1; tail -25 $OUT/pytest_conv.log | cut -c1-200
if grep -q " passed" $OUT/pytest_conv.log && ! grep -q "failed" $OUT/pytest_conv.log; then
  EGAZE_CONV_PAIR=1 timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; tail -3 $OUT/pytest_gpu.log
  for pr in 0 1; do EGAZE_CONV_PAIR=$pr timeout 300 python tools/conv_prof.py > $OUT/conv_prof_pair$pr.txt 2>&1; cut -c1-100 $OUT/conv_prof_pair$pr.txt; done
  for pr in 0 1; do EGAZE_CONV_PAIR=$pr timeout 300 python tools/layer_table.py > $OUT/layer_table_pair$pr.txt 2>&1; tail -1 $OUT/layer_table_pair$pr.txt; done
fi
