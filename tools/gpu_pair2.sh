#!/bin/bash
OUT=gpurun_out/${1:-pair2}; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; tail -3 $OUT/pytest_gpu.log
for cfg in "0 3" "1 3" "1 2"; do
  set -- $cfg
  EGAZE_CONV_WINDOW=$1 EGAZE_CONV_WINDOW_MINSB=$2 timeout 300 python tools/layer_table.py > $OUT/layer_table_w$1_sb$2.txt 2>&1
  tail -1 $OUT/layer_table_w$1_sb$2.txt
done
EGAZE_CONV_WINDOW=1 timeout 600 python -m pytest tests/test_gpu_conv.py tests/test_gpu_backward.py -m gpu -x -q > $OUT/pytest_win.log 2>&1; tail -2 $OUT/pytest_win.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $OUT/bench_train.json 2> $OUT/bench_train.err; cut -c1-200 $OUT/bench_train.json
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload sp_fwd > $OUT/bench_fwd.json 2> $OUT/bench_fwd.err; cut -c1-200 $OUT/bench_fwd.json
