#!/bin/bash
# round-2 GPU call W (1 GPU): col_attn_fa with tile 1 staggered, tight role waits: bench + timeline
mkdir -p gpurun_out
O=gpurun_out
RNAMSM_COL_IMPL=fa timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "col" > $O/r2w_col_tests_fa.log 2>&1; echo "rc=$?" >> $O/r2w_col_tests_fa.log
tail -3 $O/r2w_col_tests_fa.log
SH="512 256 4096 128 1024 1024 256 300 384 200 768 64"
: > $O/r2w_col_bench.txt
for v in "RNAMSM_COL_IMPL=fa" "RNAMSM_COL_IMPL=fa RNAMSM_COL_POLY=2"; do
  echo "== $v" >> $O/r2w_col_bench.txt
  env $v timeout 300 python tools/col_bench.py $SH >> $O/r2w_col_bench.txt 2>&1
done
cat $O/r2w_col_bench.txt
RNAMSM_COL_IMPL=fa RNAMSM_COL_TRACE=gpurun_out/r2w_trace.txt timeout 300 python tools/col_bench.py 1024 256 > gpurun_out/r2w_trace.log 2>&1
