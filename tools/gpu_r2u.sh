#!/bin/bash
# Round-2 GPU call U: what the AT + LF part costs inside the step: full_train vs sp_train in one session, LF standalone
TAG=${1:-r02u}; OUT=gpurun_out/$TAG; mkdir -p $OUT
for wl in full_train sp_train full_train sp_train; do
  timeout 600 python bench.py --workload $wl --steps 20 --warmup 5 --no-cpu-baseline --no-dropin > $OUT/bench_$wl.json 2> $OUT/bench_$wl.err
  python - <<PY
import json
d=json.load(open("$OUT/bench_$wl.json"))
print("$wl: %.1f fps  %.3f ms/step  e2e %.1f  kernels %.2f ms" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["kernel_ms_per_step"]))
PY
done
EGAZE_BENCH_LF_STREAM=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-dropin > $OUT/bench_full_nolfstream.json 2>/dev/null; python -c "
import json; d=json.load(open('$OUT/bench_full_nolfstream.json')); print('full_train, LF on the main stream: %.3f ms/step' % d['ms_per_step'])"
timeout 300 python tools/lf_bench.py 2>&1 | tail -4
