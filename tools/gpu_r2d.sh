#!/bin/bash
# Round-2 GPU call D (2 GPUs): NCCL gradient-equality test + 2-rank contract bench, overlapped vs flat all-reduce.
TAG=${1:-r02d}; OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi --query-gpu=index,name,clocks.sm --format=csv > $OUT/smi.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_ddp.py tests/test_gpu_optim.py tests/test_gpu_backward.py tests/test_gpu_lf.py -m gpu -q -s -p no:cacheprovider > $OUT/pytest.log 2>&1; echo "pytest exit $?" | tee -a $OUT/summary.txt
grep -E "passed|failed|DDP|overlap_eager" $OUT/pytest.log | tail -5; grep -E "^E  " $OUT/pytest.log | head
N=${NGPU:-2}
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $N --steps 10 --warmup 3 ${@:3} > $OUT/$2.json 2> $OUT/$2.err; tail -c 1500 $OUT/$2.json; tail -3 $OUT/$2.err; }
run 29611 bench_full_train_${N}gpu --no-dropin
EGAZE_BENCH_DDP=flat run 29612 bench_full_train_${N}gpu_flat --no-dropin
run 29613 bench_sp_train_${N}gpu --workload sp_train --no-dropin
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-dropin > $OUT/bench_full_train_1gpu.json 2> $OUT/bench_full_train_1gpu.err; tail -c 400 $OUT/bench_full_train_1gpu.json
ls -la $OUT
