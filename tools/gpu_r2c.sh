#!/bin/bash
# Round-2 GPU call C: full GPU suite (per file, own process) + contract bench (fused Adam) + A/B vs torch Adam.
TAG=${1:-r02c}; OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
for f in test_gpu_optim test_gpu_backward test_gpu_models test_gpu_lf test_gpu_conv test_gpu_small test_gpu_golden test_gpu_graph test_gpu_ddp; do
  timeout 900 python -m pytest tests/$f.py -m gpu -q -s -p no:cacheprovider > $OUT/$f.log 2>&1; echo "$f exit $?" | tee -a $OUT/summary.txt
  grep -E "passed|failed|error" $OUT/$f.log | tail -2
  grep -E "^E  " $OUT/$f.log | head -6
done
grep -h -E "B=32x224|eval-mode BatchNorm|egaze-vs-fp64" $OUT/test_gpu_models.log $OUT/test_gpu_backward.log | head
timeout 600 python bench.py --steps 10 --warmup 3 > $OUT/bench_full_train.json 2> $OUT/bench_full_train.err; tail -c 2800 $OUT/bench_full_train.json; tail -5 $OUT/bench_full_train.err
EGAZE_BENCH_ADAM=torch timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-dropin > $OUT/bench_full_train_torchadam.json 2> $OUT/bench_full_train_torchadam.err; tail -c 700 $OUT/bench_full_train_torchadam.json
timeout 600 python bench.py --steps 10 --warmup 3 --workload sp_train --no-cpu-baseline > $OUT/bench_sp_train.json 2> $OUT/bench_sp_train.err; tail -c 900 $OUT/bench_sp_train.json; tail -3 $OUT/bench_sp_train.err
ls -la $OUT
