#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_backward.py tests/test_gpu_conv.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3
timeout 300 python -m pytest tests/test_gpu_models.py tests/test_gpu_golden.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -2
bash tools/gpu_ab_lib.sh ${1:-ab}
EGAZE_WGRAD_PAIRS=1 timeout 300 python tools/layer_table.py 2>&1 | grep -E "wgrad" | awk '{print $1,$3,$4,$5,$8}' | sort | uniq -c | sort -k2 | head -20
