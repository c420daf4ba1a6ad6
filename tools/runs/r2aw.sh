#!/bin/bash
# round-2 GPU call AW (1 GPU): cost of the GELU epilogue (fc1 with bias-only vs bias + GELU)
mkdir -p gpurun_out
timeout 200 python tools/epi_cost_bench.py 131072 > gpurun_out/r2aw_epi_cost.txt 2>&1
timeout 200 python tools/epi_cost_bench.py 18432 >> gpurun_out/r2aw_epi_cost.txt 2>&1
cat gpurun_out/r2aw_epi_cost.txt
