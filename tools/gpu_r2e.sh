#!/bin/bash
# Round-2 GPU call E: profiler evidence.  ncu launch list of the contract bench command (graph replay), `ncu --set full` of the
# representative layers + LF kernels, compute-sanitizer memcheck / racecheck over small-shape tests.
TAG=${1:-r02e}; OUT=gpurun_out/$TAG; mkdir -p $OUT
if [ -z "$SKIP_LIST" ]; then
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file $OUT/launches_full_train.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-dropin > $OUT/bench_under_ncu.log 2>&1
python tools/launch_summary.py $OUT/launches_full_train.csv lf_head_fwd_kernel 45 > $OUT/launch_summary_full_train.txt 2>&1; head -50 $OUT/launch_summary_full_train.txt
fi
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'conv3x3_tc|wgrad_tc|lf_conv|lf_wgrad' -c 20 -f -o $OUT/prof_layers \
  python tools/ncu_conv.py > $OUT/ncu_conv.log 2>&1; tail -12 $OUT/ncu_conv.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_conv.py tests/test_gpu_lf.py tests/test_gpu_optim.py -m gpu -q -x -k "not 224 and not shape5 and not 112" -p no:cacheprovider > $OUT/sanitizer_memcheck.log 2>&1; echo "memcheck exit $?" | tee -a $OUT/summary.txt; grep -n -B2 -A25 'Invalid\|Error:\|error:' $OUT/sanitizer_memcheck.log | head -80; tail -5 $OUT/sanitizer_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_gpu_lf.py -m gpu -q -x -k "shape0 or 32-48" -p no:cacheprovider > $OUT/sanitizer_racecheck_lf.log 2>&1; echo "racecheck lf exit $?" | tee -a $OUT/summary.txt; tail -5 $OUT/sanitizer_racecheck_lf.log
ls -la $OUT
