#!/bin/bash
# round-2 GPU call N (1 GPU): tcgen05.mma issue / dependency microbenchmark
mkdir -p gpurun_out
timeout 120 tools/micro/mma_chain_bench > gpurun_out/r2n_mma_chain.txt 2>&1; echo "rc=$?" >> gpurun_out/r2n_mma_chain.txt
cat gpurun_out/r2n_mma_chain.txt
