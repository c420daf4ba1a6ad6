#!/bin/bash
# round-2 GPU call P (1 GPU): ncu capture of the new column attention (stall sampling by SASS line)
mkdir -p gpurun_out
O=gpurun_out
RNAMSM_COL_IMPL=fa timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:col_attn_fa_kernel' -s 5 -c 1 -o $O/r2r_prof_col_fa python tools/col_bench.py 1024 256 > $O/r2r_ncu.log 2>&1
tail -3 $O/r2r_ncu.log; ls -la $O/*.ncu-rep
