#!/bin/bash
# A/B of two builds of libegaze.so inside ONE GPU session (boxes differ by a few per cent, runs inside a session by < 1 %):
# egaze/libegaze_A.so (baseline) against egaze/libegaze.so, the contract bench alternating twice.
TAG=${1:-ab}; OUT=gpurun_out/$TAG; mkdir -p $OUT
A=$PWD/egocentric-gaze-prediction_b200/egaze/libegaze_A.so
for rep in 1 2; do
  for v in A B; do
    if [ $v = A ]; then export EGAZE_LIB=$A; else unset EGAZE_LIB; fi
    timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-dropin > $OUT/bench_${v}_$rep.json 2> $OUT/bench_${v}_$rep.err
    python - <<PY
import json
d=json.load(open("$OUT/bench_${v}_$rep.json"))
print("$v rep $rep: %.1f fps  %.3f ms/step  e2e %.1f  kernels %.2f ms  frac %.3f  sm %s MHz" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["kernel_ms_per_step"], d["roofline"]["frac"], d["clocks"]["sm_mhz"]))
PY
  done
done
