#!/bin/bash
# conv tests on the new build, then the same-session A/B of two builds
timeout 600 python -m pytest tests/test_gpu_conv.py tests/test_gpu_golden.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3
bash tools/gpu_ab_lib.sh ${1:-ab}
