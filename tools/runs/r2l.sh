#!/bin/bash
# round-2 GPU call L (1 GPU): resource microbenchmark for the column attention
mkdir -p gpurun_out
O=gpurun_out
timeout 120 tools/micro/tmem_umma_bench > $O/r2l_tmem_umma2.txt 2>&1; echo "rc=$?" >> $O/r2l_tmem_umma2.txt
cat $O/r2l_tmem_umma2.txt
