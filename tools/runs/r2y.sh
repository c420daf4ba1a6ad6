#!/bin/bash
# round-2 GPU call Y (1 GPU): col_attn_fa (one warpgroup per tile), knobs A/B
mkdir -p gpurun_out
O=gpurun_out
RNAMSM_COL_IMPL=fa timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "col" > $O/r2y_col_tests_fa.log 2>&1; echo "rc=$?" >> $O/r2y_col_tests_fa.log
tail -3 $O/r2y_col_tests_fa.log
SH="512 256 4096 128 1024 1024 256 300"
: > $O/r2y_col_bench.txt
for v in "RNAMSM_COL_IMPL=fa" "RNAMSM_COL_IMPL=fa RNAMSM_COL_TIGHT=0" "RNAMSM_COL_IMPL=fa RNAMSM_COL_STAGGER=1300" "RNAMSM_COL_IMPL=fa RNAMSM_COL_POLY=2" "RNAMSM_COL_IMPL=fa RNAMSM_COL_POLY=4"; do
  echo "== $v" >> $O/r2y_col_bench.txt
  env $v timeout 300 python tools/col_bench.py $SH >> $O/r2y_col_bench.txt 2>&1
done
cat $O/r2y_col_bench.txt
