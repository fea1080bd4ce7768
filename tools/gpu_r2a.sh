#!/bin/bash
# Round-2 GPU call A: new LF kernels (tests + timing), full GPU suite, the contract bench (full_train, graph-captured) and
# sp_train beside it, backward-precision experiment.      usage: tools/gpu_r2a.sh <tag>
TAG=${1:-r02a}; OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_lf.py -m gpu -q -s > $OUT/pytest_lf.log 2>&1; echo "pytest_lf exit $?" | tee -a $OUT/pytest_lf.log
tail -15 $OUT/pytest_lf.log
timeout 300 python tools/lf_bench.py > $OUT/lf_bench.txt 2>&1; cat $OUT/lf_bench.txt | tail -6
timeout 1500 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a $OUT/pytest_gpu.log
tail -25 $OUT/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 > $OUT/bench_full_train.json 2> $OUT/bench_full_train.err; tail -c 2500 $OUT/bench_full_train.json; tail -5 $OUT/bench_full_train.err
timeout 600 python bench.py --steps 10 --warmup 3 --workload sp_train --no-cpu-baseline > $OUT/bench_sp_train.json 2> $OUT/bench_sp_train.err; tail -c 1800 $OUT/bench_sp_train.json; tail -3 $OUT/bench_sp_train.err
timeout 900 python tools/grad_modes.py > $OUT/grad_modes.txt 2>&1; cat $OUT/grad_modes.txt | tail -40
ls -la $OUT
