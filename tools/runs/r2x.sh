#!/bin/bash
# round-2 GPU call X (1 GPU): col_attn_fa with two softmax warpgroups per tile (column split)
mkdir -p gpurun_out
O=gpurun_out
RNAMSM_COL_IMPL=fa timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "col" > $O/r2x_col_tests_fa.log 2>&1; echo "rc=$?" >> $O/r2x_col_tests_fa.log
tail -8 $O/r2x_col_tests_fa.log
SH="512 256 4096 128 1024 1024 256 300 384 200 768 64"
: > $O/r2x_col_bench.txt
for v in "RNAMSM_COL_IMPL=fa" "RNAMSM_COL_IMPL=fa RNAMSM_COL_POLY=2"; do
  echo "== $v" >> $O/r2x_col_bench.txt
  env $v timeout 300 python tools/col_bench.py $SH >> $O/r2x_col_bench.txt 2>&1
done
cat $O/r2x_col_bench.txt
RNAMSM_COL_IMPL=fa timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:col_attn_fa_kernel' -s 5 -c 1 -o $O/r2x_prof_col_fa python tools/col_bench.py 1024 256 > $O/r2x_ncu.log 2>&1
