#!/bin/bash
# N-GPU A/B of NCCL's SM footprint under the overlapped gradient all-reduce (persistent conv kernels own every SM: the CTAs NCCL takes
# delay whole waves of conv tiles).  usage: tools/gpu_nccl_ab.sh <tag> <N>
TAG=${1:-nccl}; N=${2:-4}; OUT=gpurun_out/$TAG; mkdir -p $OUT
run() { timeout 600 env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 20 --warmup 5 > $OUT/b.json 2> $OUT/b.err; python -c "
import json; d=json.load(open('$OUT/b.json')); print('%-40s %d GPUs  %.1f fps  %.3f ms/step' % ('$*', d['n_gpus'], d['value'], d['ms_per_step']))" || tail -3 $OUT/b.err; }
run A=1
run NCCL_MAX_CTAS=8
run NCCL_MAX_CTAS=4
run NCCL_MAX_CTAS=16
run A=1
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-dropin > $OUT/b1.json 2>/dev/null; python -c "
import json; d=json.load(open('$OUT/b1.json')); print('1 GPU (rank 0 GPU): %.1f fps  %.3f ms/step' % (d['value'], d['ms_per_step']))"
