#!/bin/bash
# round-2 GPU call BC (1 GPU): A V GEMM with the 16-warp epilogue
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "row_attention" > $O/r2bc_tests.log 2>&1; echo "rc=$?" >> $O/r2bc_tests.log; tail -3 $O/r2bc_tests.log
for wv in 1 0; do
  for shp in "512 256" "256 300" "1024 1024"; do
    RNAMSM_AV_WIDE=$wv GEMM_BENCH_ONE=1 timeout 200 python tools/gemm_bench.py $shp 2>&1 | grep -E "row_av|row_logits" | sed "s/^/AV_WIDE=$wv /"
  done
done | tee $O/r2bc_av.txt
