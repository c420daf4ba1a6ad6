#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
for dbg in 0 1 2 4 8 6 12; do
  echo "RNAMSM_LN_DEBUG=$dbg" >> $O/r2e_lnfuse_dbg.txt
  RNAMSM_LN_DEBUG=$dbg timeout 200 python tools/lnfuse_bench.py >> $O/r2e_lnfuse_dbg.txt 2>&1
done
cat $O/r2e_lnfuse_dbg.txt
