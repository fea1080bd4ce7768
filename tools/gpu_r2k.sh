#!/bin/bash
# Round-2 GPU call K: A/B of the MMA warp's tap loop (EGAZE_CONV_PROBE = 1: round-1 loop with fused barrier probes, 0: lean unrolled
# loop) on the per-layer table of one full train step, conv kernel tests under both.
TAG=${1:-r02k}; OUT=gpurun_out/$TAG; mkdir -p $OUT
for pb in ${PROBES:-0 1}; do
  EGAZE_CONV_PROBE=$pb timeout 300 python tools/layer_table.py > $OUT/layer_table_probe$pb.txt 2>&1
  echo "PROBE=$pb: $(tail -1 $OUT/layer_table_probe$pb.txt)"
  grep -E "224x224 +64->64|224x224 +16->64|112x112 128->64|112x112 +64->128 red=0 ups=0 stats=1|56x56 +256->256|28x28 +512->512" $OUT/layer_table_probe$pb.txt | awk '{print $1,$3,$4,$8,$9,$10,$11}' | sort | uniq -c | sort -k2 | head -40
done
timeout 600 python -m pytest tests/test_gpu_conv.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -2
