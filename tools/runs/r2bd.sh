#!/bin/bash
# round-2 GPU call BD (1 GPU): QKV (bias-only epilogue) on the 16-warp / 5-stage instance?
mkdir -p gpurun_out
for wv in 0 1 0 1; do
  RNAMSM_DENSE_WIDE=$wv timeout 200 python tools/epi_cost_bench.py 131072 2>&1 | grep "bias only" | sed "s/^/DENSE_WIDE=$wv /"
done | tee gpurun_out/r2bd_dense_wide.txt
