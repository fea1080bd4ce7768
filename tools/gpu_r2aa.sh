#!/bin/bash
# Round-2 GPU call AA: reference arm at the workload's own batch (host memory permitting), BASELINE configs[4] sweep on one GPU
TAG=${1:-r02aa}; OUT=gpurun_out/$TAG; mkdir -p $OUT
free -g | head -2; nproc
/usr/bin/time -v timeout 900 python bench.py --impl reference --steps 20 --warmup 3 > $OUT/bench_reference.json 2> $OUT/ref.err; cut -c1-700 $OUT/bench_reference.json; grep -E "Maximum resident|Elapsed" $OUT/ref.err
timeout 1200 python tools/sweep.py > $OUT/sweep_1gpu.jsonl 2> $OUT/sweep.err; cat $OUT/sweep_1gpu.jsonl | cut -c1-250
