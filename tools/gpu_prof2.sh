#!/bin/bash
OUT=gpurun_out/${1:-prof2}; mkdir -p $OUT
run() { tag=$1; shift; env "$@" timeout 300 python tools/conv_prof.py > $OUT/conv_prof_$tag.txt 2>&1; echo "== $tag"; cut -c1-100 $OUT/conv_prof_$tag.txt; }
run grouped EGAZE_CONV_GROUPED=1
run unmerged EGAZE_CONV_MERGED=0
run fast EGAZE_PRECISION=fast
for st in 0 1; do EGAZE_WGRAD_STACKED=$st timeout 300 python tools/layer_table.py > $OUT/layer_table_stacked$st.txt 2>&1; tail -1 $OUT/layer_table_stacked$st.txt; done
