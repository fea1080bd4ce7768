#!/bin/bash
# Round-2 GPU call X (N GPUs): the 2-rank NCCL gradient-equality test, then the contract bench under torchrun
N=${2:-2}; TAG=${1:-r02x}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_ddp.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 20 --warmup 5 > $OUT/bench_full_train_${N}gpu.json 2> $OUT/bench_${N}gpu.err; tail -c 2500 $OUT/bench_full_train_${N}gpu.json; tail -3 $OUT/bench_${N}gpu.err
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-dropin > $OUT/bench_full_train_1gpu.json 2>/dev/null; python -c "
import json; a=json.load(open('$OUT/bench_full_train_1gpu.json')); b=json.load(open('$OUT/bench_full_train_${N}gpu.json'))
print('1 GPU %.1f fps %.3f ms | $N GPUs %.1f fps %.3f ms | efficiency %.3f' % (a['value'], a['ms_per_step'], b['value'], b['ms_per_step'], b['value']/($N*a['value'])))"
