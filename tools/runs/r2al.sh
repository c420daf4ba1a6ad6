#!/bin/bash
# round-2 GPU call AL (1 GPU): phase timeline of row_attn_short (CTA 0) at four shapes
mkdir -p gpurun_out
RNAMSM_SHORT_TRACE=1 timeout 300 python tools/row_short_bench.py 512 36 256 64 256 100 4096 128 2>&1 | grep -v "^$" > gpurun_out/r2al_trace.txt
sort gpurun_out/r2al_trace.txt | uniq -c | sort -k3,3 -k5,5 | awk '{print}' | tail -60
