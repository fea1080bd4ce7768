#!/bin/bash
OUT=gpurun_out/${1:-prod}; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; tail -3 $OUT/pytest_gpu.log
timeout 300 python tools/conv_prof.py > $OUT/conv_prof.txt 2>&1; cut -c1-100 $OUT/conv_prof.txt; cut -c1-32,100-330 $OUT/conv_prof.txt | head -5
timeout 300 python tools/layer_table.py > $OUT/layer_table.txt 2>&1; tail -1 $OUT/layer_table.txt
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $OUT/bench_train.json 2> $OUT/bench_train.err; cut -c1-200 $OUT/bench_train.json
