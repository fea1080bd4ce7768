#!/bin/bash
# Round-2 GPU call I: validate HEAD after the container re-creation -- whole GPU suite (one process per file), the contract
# bench (default arguments, exactly what the driver runs), the per-layer table, the drop-in loop.
TAG=${1:-r02i}; OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
for t in conv backward models golden graph optim small lf data; do
  timeout 900 python -m pytest tests/test_gpu_$t.py -m gpu -q -x -p no:cacheprovider > $OUT/test_gpu_$t.log 2>&1; echo "test_gpu_$t exit $?" | tee -a $OUT/summary.txt
  tail -2 $OUT/test_gpu_$t.log | cut -c1-200; grep -E "^E  " $OUT/test_gpu_$t.log | head -8 | cut -c1-220
done
timeout 600 python bench.py > $OUT/bench_full_train.json 2> $OUT/bench_full_train.err; tail -c 3000 $OUT/bench_full_train.json; tail -3 $OUT/bench_full_train.err
timeout 300 python tools/layer_table.py > $OUT/layer_table.txt 2>&1; tail -1 $OUT/layer_table.txt
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
ls -la $OUT
