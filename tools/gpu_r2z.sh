#!/bin/bash
# Round-2 GPU call Z: B = 64 through the graphed bench (is the sweep's B = 64 point a host / allocator artefact of its eager loop?),
# and what the side streams buy (EGAZE_TRUNK_STREAM / EGAZE_WGRAD_STREAM off)
TAG=${1:-r02z}; OUT=gpurun_out/$TAG; mkdir -p $OUT
run() { timeout 600 env "$@" python bench.py --workload sp_train --steps 10 --warmup 4 --no-cpu-baseline --no-dropin ${EXTRA} > $OUT/b.json 2>/dev/null; python -c "
import json; d=json.load(open('$OUT/b.json')); print('%-48s %.1f fps  %.3f ms/step  e2e %.1f  kernels %.2f ms  host %.2f ms' % ('$* $EXTRA', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['kernel_ms_per_step'], d.get('host_enqueue_ms_per_step', 0)))"; }
EXTRA="--batch 64" run A=1
EXTRA="--batch 64" run EGAZE_BENCH_GRAPH=0
EXTRA="" run EGAZE_BENCH_GRAPH=0
EXTRA="" run A=1
EXTRA="" run EGAZE_TRUNK_STREAM=0
EXTRA="" run EGAZE_WGRAD_STREAM=0
EXTRA="" run EGAZE_TRUNK_STREAM=0 EGAZE_WGRAD_STREAM=0
EXTRA="" run A=1
