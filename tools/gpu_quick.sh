#!/bin/bash
# quick check: full GPU test-suite + device-timed train step / layer table   usage: tools/gpu_quick.sh <tag>
OUT=gpurun_out/${1:-quick}; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; tail -3 $OUT/pytest_gpu.log
timeout 300 python tools/layer_table.py > $OUT/layer_table.txt 2>&1; tail -1 $OUT/layer_table.txt
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $OUT/bench_train.json 2> $OUT/bench_train.err; cut -c1-400 $OUT/bench_train.json
