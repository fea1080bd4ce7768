#!/bin/bash
# Round-2 GPU call B: fp16-split forward + cheaper backward (new default numeric mode).  usage: tools/gpu_r2b.sh <tag>
TAG=${1:-r02b}; OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
for f in test_gpu_conv test_gpu_small test_gpu_backward test_gpu_models test_gpu_golden test_gpu_graph test_gpu_lf; do
  timeout 900 python -m pytest tests/$f.py -m gpu -q -s -p no:cacheprovider > $OUT/$f.log 2>&1; echo "$f exit $?" | tee -a $OUT/summary.txt
  grep -E "passed|failed|error" $OUT/$f.log | tail -2
done
grep -h -E "max-abs|rel-L2|B=32x224|worst" $OUT/test_gpu_models.log $OUT/test_gpu_golden.log $OUT/test_gpu_backward.log | head -40
timeout 600 python tools/grad_modes.py > $OUT/grad_modes.txt 2>&1; head -8 $OUT/grad_modes.txt
timeout 600 python bench.py --steps 10 --warmup 3 > $OUT/bench_full_train.json 2> $OUT/bench_full_train.err; tail -c 2600 $OUT/bench_full_train.json; tail -5 $OUT/bench_full_train.err
timeout 600 python bench.py --steps 10 --warmup 3 --workload sp_train --no-cpu-baseline > $OUT/bench_sp_train.json 2> $OUT/bench_sp_train.err; tail -c 1200 $OUT/bench_sp_train.json; tail -3 $OUT/bench_sp_train.err
EGAZE_PRECISION=precise3 timeout 600 python bench.py --steps 10 --warmup 3 --workload sp_train --no-cpu-baseline --no-dropin > $OUT/bench_sp_train_precise3.json 2> $OUT/bench_sp_train_precise3.err; tail -c 600 $OUT/bench_sp_train_precise3.json
timeout 300 python tools/layer_table.py > $OUT/layer_table.txt 2>&1; tail -3 $OUT/layer_table.txt
ls -la $OUT
