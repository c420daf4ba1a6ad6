#!/bin/bash
# round-2 GPU call O (1 GPU): the 128-key-step column attention with P in tensor memory (RNAMSM_COL_IMPL=fa)
mkdir -p gpurun_out
O=gpurun_out
RNAMSM_COL_IMPL=fa timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "col" > $O/r2o_col_tests_fa.log 2>&1; echo "rc=$?" >> $O/r2o_col_tests_fa.log
tail -15 $O/r2o_col_tests_fa.log
SH="512 256 4096 128 1024 1024 256 300 384 200 768 64"
: > $O/r2o_col_bench.txt
for v in "RNAMSM_COL_IMPL=ws" "RNAMSM_COL_IMPL=ws RNAMSM_COL_GROUPS=2" "RNAMSM_COL_IMPL=fa" "RNAMSM_COL_IMPL=fa RNAMSM_COL_POLY=4" "RNAMSM_COL_IMPL=fa RNAMSM_COL_POLY=2" "RNAMSM_COL_IMPL=fa RNAMSM_COL_POLY=1"; do
  echo "== $v" >> $O/r2o_col_bench.txt
  env $v timeout 300 python tools/col_bench.py $SH >> $O/r2o_col_bench.txt 2>&1
done
cat $O/r2o_col_bench.txt
