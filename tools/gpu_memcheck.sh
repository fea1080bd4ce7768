OUT=gpurun_out/r02e2; mkdir -p $OUT
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_conv.py tests/test_gpu_lf.py tests/test_gpu_optim.py -m gpu -q -x -k "not 224 and not shape5 and not 112" -p no:cacheprovider > $OUT/sanitizer_memcheck.log 2>&1; echo "memcheck exit $?"
grep -n "=========" $OUT/sanitizer_memcheck.log | head -60
