#!/bin/bash
TAG=${1:-r02f}; OUT=gpurun_out/$TAG; mkdir -p $OUT
for f in test_gpu_data test_gpu_conv test_gpu_optim test_gpu_backward test_gpu_models test_gpu_lf test_gpu_small test_gpu_golden test_gpu_graph; do
  timeout 900 python -m pytest tests/$f.py -m gpu -q -p no:cacheprovider > $OUT/$f.log 2>&1; echo "$f exit $?" | tee -a $OUT/summary.txt
  grep -E "passed|failed|error" $OUT/$f.log | tail -1; grep -E "^E  " $OUT/$f.log | head -6
done
timeout 300 python tools/lf_bench.py > $OUT/lf_bench.txt 2>&1; head -2 $OUT/lf_bench.txt
timeout 600 python bench.py --steps 10 --warmup 3 > $OUT/bench_full_train.json 2> $OUT/bench_full_train.err; tail -c 1500 $OUT/bench_full_train.json; tail -5 $OUT/bench_full_train.err
EGAZE_CONV_WSTAT=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-dropin > $OUT/bench_full_train_nowstat.json 2> $OUT/bench_full_train_nowstat.err; tail -c 400 $OUT/bench_full_train_nowstat.json
timeout 300 python tools/layer_table.py > $OUT/layer_table.txt 2>&1; tail -2 $OUT/layer_table.txt
