#!/bin/bash
# Round-2 GPU call Y: final validation of HEAD -- whole GPU suite, the contract bench with default arguments, per-layer table, the
# reference arm, smoke(), BASELINE configs[4] sweep on one GPU
TAG=${1:-r02y}; OUT=gpurun_out/$TAG; mkdir -p $OUT
for t in conv backward models golden graph optim small lf data; do
  timeout 900 python -m pytest tests/test_gpu_$t.py -m gpu -q -x -p no:cacheprovider > $OUT/test_gpu_$t.log 2>&1; echo "test_gpu_$t exit $?" | tee -a $OUT/summary.txt
  tail -1 $OUT/test_gpu_$t.log | cut -c1-200; grep -E "^E  " $OUT/test_gpu_$t.log | head -8 | cut -c1-220
done
timeout 600 python bench.py > $OUT/bench_full_train.json 2> $OUT/bench_full_train.err; tail -c 600 $OUT/bench_full_train.json; tail -3 $OUT/bench_full_train.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_reference.json 2>/dev/null; cut -c1-400 $OUT/bench_reference.json
timeout 300 python tools/layer_table.py > $OUT/layer_table.txt 2>&1; tail -1 $OUT/layer_table.txt
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
for wl in sp_train sp_fwd pipeline_fwd at_seq; do timeout 300 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline --no-dropin > $OUT/bench_$wl.json 2>/dev/null; python -c "
import json; d=json.load(open('$OUT/bench_$wl.json')); print('$wl: %.1f %s  %.3f ms/step  e2e %.1f' % (d['value'], d['unit'], d['ms_per_step'], d['e2e']['value']))"; done
timeout 900 python tools/sweep.py > $OUT/sweep_1gpu.jsonl 2> $OUT/sweep.err; tail -16 $OUT/sweep_1gpu.jsonl | cut -c1-250
