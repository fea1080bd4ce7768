#!/bin/bash
# Exploratory GPU run: each test file in its own process (a trapped kernel poisons the CUDA context), bounded by timeout.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
for f in "$@"; do
  name=$(basename $f .py)
  timeout 600 python -m pytest $f -m gpu -q --no-header -p no:cacheprovider > gpurun_out/$name.log 2>&1
  echo "$name exit=$?" | tee -a gpurun_out/summary.txt
  tail -n 30 gpurun_out/$name.log
done
