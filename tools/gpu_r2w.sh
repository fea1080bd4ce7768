#!/bin/bash
# Round-2 GPU call W: final profiler evidence (sized for gpurun's 64 MiB return limit): ncu launch list of the contract bench command
# (graph replay), per-launch DRAM traffic of the tcgen05 launches, `ncu --set full` of representative layers exported to CSV on the
# box, compute-sanitizer memcheck / racecheck over the kernel tests.
TAG=${1:-r02w}; OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file $OUT/launches_full_train.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-dropin > $OUT/bench_under_ncu.log 2>&1
python tools/launch_summary.py $OUT/launches_full_train.csv lf_head_fwd_kernel 60 > $OUT/launch_summary_full_train.txt 2>&1; head -50 $OUT/launch_summary_full_train.txt
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:'conv3x3_tc|wgrad_tc' -c 4000 --csv \
  --log-file $OUT/conv_traffic.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-dropin > $OUT/bench_under_ncu2.log 2>&1
python tools/traffic_summary.py $OUT/conv_traffic.csv full_train
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'conv3x3_tc|wgrad_tc|lf_conv|lf_wgrad|lf_dgrad' -c 30 -f -o $OUT/prof_layers \
  python tools/ncu_conv.py > $OUT/ncu_conv.log 2>&1; grep -v PROF $OUT/ncu_conv.log | tail -14
ncu -i $OUT/prof_layers.ncu-rep --page raw --csv > $OUT/prof_layers_raw.csv 2>/dev/null
python tools/ncu_summary.py $OUT/prof_layers_raw.csv > $OUT/ncu_full_layers_summary.csv; cut -c1-200 $OUT/ncu_full_layers_summary.csv | head -32
for id in 0 10; do ncu -i $OUT/prof_layers.ncu-rep --page source --csv --launch-skip $id --launch-count 1 > $OUT/prof_source_launch$id.csv 2>/dev/null; done
gzip -f $OUT/prof_source_launch*.csv $OUT/prof_layers_raw.csv $OUT/launches_full_train.csv $OUT/conv_traffic.csv
rm -f $OUT/prof_layers.ncu-rep
timeout 900 compute-sanitizer --tool memcheck --report-api-errors no --error-exitcode 7 python -m pytest tests/test_gpu_conv.py tests/test_gpu_lf.py tests/test_gpu_optim.py tests/test_gpu_small.py -m gpu -q -x -k "not 224 and not shape5 and not 112" -p no:cacheprovider > $OUT/sanitizer_memcheck.log 2>&1; echo "memcheck exit $?" | tee -a $OUT/summary.txt; tail -4 $OUT/sanitizer_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_gpu_lf.py tests/test_gpu_conv.py -m gpu -q -x -k "shape0 or 32-48 or plain" -p no:cacheprovider > $OUT/sanitizer_racecheck.log 2>&1; echo "racecheck exit $?" | tee -a $OUT/summary.txt; tail -4 $OUT/sanitizer_racecheck.log
du -sh $OUT; ls -la $OUT
