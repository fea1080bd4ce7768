#!/bin/bash
# One gpurun call: GPU parity tests, contract bench, ncu launch list of the bench command, DRAM traffic of every conv launch,
# ncu --set full of the hot kernels.     usage: tools/gpu_round.sh <tag> [tests|notests]
TAG=${1:-rXX}; MODE=${2:-tests}
OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
if [ "$MODE" = "tests" ]; then
  timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log
  tail -5 $OUT/pytest_gpu.log
fi
timeout 600 python bench.py --steps 10 --warmup 3 > $OUT/bench_train.json 2> $OUT/bench_train.err; tail -c 1500 $OUT/bench_train.json
timeout 300 python bench.py --steps 10 --warmup 3 --workload sp_fwd --no-cpu-baseline > $OUT/bench_fwd.json 2> $OUT/bench_fwd.err; tail -c 600 $OUT/bench_fwd.json
timeout 300 python tools/layer_table.py > $OUT/layer_table.txt 2>&1; tail -1 $OUT/layer_table.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file $OUT/launches_train.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/bench_under_ncu.log 2>&1
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:'conv3x3_tc|wgrad_tc' -c 4000 --csv \
  --log-file $OUT/conv_traffic.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/bench_under_ncu2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'conv3x3_tc|wgrad_tc' -c 12 -f -o $OUT/prof_conv \
  python tools/ncu_conv.py > $OUT/ncu_conv.log 2>&1; tail -4 $OUT/ncu_conv.log
ls -la $OUT
