#!/bin/bash
# same-session A/B of environment settings on the contract bench: tools/gpu_env_ab.sh <tag> "VAR=a" "VAR=b" ...
TAG=$1; shift; OUT=gpurun_out/$TAG; mkdir -p $OUT
for rep in 1 2; do for cfg in "$@"; do
  timeout 600 env $cfg python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-dropin > $OUT/b.json 2>/dev/null; python -c "
import json; d=json.load(open('$OUT/b.json')); print('%-28s rep $rep: %.1f fps  %.3f ms/step  kernels %.2f ms  sm %s MHz' % ('$cfg', d['value'], d['ms_per_step'], d['roofline']['kernel_ms_per_step'], d['clocks']['sm_mhz']))"
done; done
