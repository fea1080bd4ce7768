#!/bin/bash
OUT=gpurun_out/${1:-prof}; mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_backward.py tests/test_gpu_conv.py -q -m gpu -x > $OUT/pytest.log 2>&1; tail -2 $OUT/pytest.log
timeout 300 python tools/conv_prof.py > $OUT/conv_prof_w0.txt 2>&1; cat $OUT/conv_prof_w0.txt
EGAZE_CONV_WINDOW=1 EGAZE_CONV_WINDOW_MINSB=2 timeout 300 python tools/conv_prof.py > $OUT/conv_prof_w1.txt 2>&1; cat $OUT/conv_prof_w1.txt
EGAZE_CONV_CLUSTER=1 timeout 300 python tools/conv_prof.py > $OUT/conv_prof_cs1.txt 2>&1; cat $OUT/conv_prof_cs1.txt
timeout 300 python tools/layer_table.py > $OUT/layer_table.txt 2>&1; tail -1 $OUT/layer_table.txt
