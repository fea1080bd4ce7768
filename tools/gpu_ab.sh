#!/bin/bash
# A/B of an environment switch on the device-timed train step (same box, alternating runs).  usage: gpu_ab.sh <tag> VAR
OUT=gpurun_out/${1:-ab}; VAR=$2; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; tail -3 $OUT/pytest_gpu.log
for rep in 1 2; do for v in 0 1; do
  env $VAR=$v timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $OUT/bench_${v}_$rep.json 2> $OUT/bench_${v}_$rep.err
  python -c "import json,sys; d=json.load(open('$OUT/bench_${v}_$rep.json')); print('$VAR=$v rep $rep: %.2f ms/step %.1f fps e2e %.1f' % (d['ms_per_step'], d['value'], d['e2e']['value']))" 2>/dev/null || tail -2 $OUT/bench_${v}_$rep.err
done; done
