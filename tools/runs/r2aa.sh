#!/bin/bash
# round-2 GPU call AA (1 GPU): in-kernel timeline of col_attn_fa (CTA 0)
mkdir -p gpurun_out
RNAMSM_COL_IMPL=fa RNAMSM_COL_TRACE=gpurun_out/r2aa_trace.txt timeout 300 python tools/col_bench.py 1024 256 > gpurun_out/r2v.log 2>&1
tail -2 gpurun_out/r2v.log; wc -l gpurun_out/r2aa_trace.txt
