#!/bin/bash
# Round-2 GPU call H: sub-pixel decoder convs -- kernel tests first, then the whole suite, the contract bench with / without
# EGAZE_SUBPIXEL, the per-layer table.
TAG=${1:-r02h}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_conv.py -m gpu -q -x -k "subpixel or planar or roundtrip" -p no:cacheprovider > $OUT/test_sub.log 2>&1; rc=$?; echo "sub tests exit $rc" | tee -a $OUT/summary.txt
tail -25 $OUT/test_sub.log | cut -c1-220
if [ $rc -ne 0 ] && [ -z "$FORCE" ]; then
  timeout 600 python -m pytest tests/test_gpu_conv.py -m gpu -q -k "subpixel or planar or roundtrip" -p no:cacheprovider 2>&1 | grep -E "^E  |passed|failed|FAILED" | head -40 | cut -c1-220
  exit 0
fi
for t in conv backward models golden graph optim small lf data; do
  timeout 900 python -m pytest tests/test_gpu_$t.py -m gpu -q -x -p no:cacheprovider > $OUT/test_gpu_$t.log 2>&1; echo "test_gpu_$t exit $?" | tee -a $OUT/summary.txt
  tail -2 $OUT/test_gpu_$t.log | cut -c1-200; grep -E "^E  " $OUT/test_gpu_$t.log | head -8 | cut -c1-220
done
timeout 600 python bench.py --steps 10 --warmup 3 --no-dropin > $OUT/bench_full_train.json 2> $OUT/bench_full_train.err; tail -c 2200 $OUT/bench_full_train.json; tail -3 $OUT/bench_full_train.err
EGAZE_SUBPIXEL=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-dropin --no-cpu-baseline > $OUT/bench_full_train_nosub.json 2> $OUT/bench_full_train_nosub.err; tail -c 1400 $OUT/bench_full_train_nosub.json
timeout 300 python tools/layer_table.py > $OUT/layer_table.txt 2>&1; tail -1 $OUT/layer_table.txt
ls -la $OUT
