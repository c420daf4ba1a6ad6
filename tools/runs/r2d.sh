#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python tools/lnfuse_bench.py > $O/r2d_lnfuse_bench.txt 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:umma_gemm_kernel -s 3 -c 1 -o $O/r2d_prof_lnfuse python tools/lnfuse_bench.py 131072 fused > $O/r2d_ncu.log 2>&1
cat $O/r2d_lnfuse_bench.txt; tail -3 $O/r2d_ncu.log
