#!/bin/bash
# round-2 GPU call AM (1 GPU): column attention at the SUSTAINED clock with a share of the exponentials on the FMA pipe
mkdir -p gpurun_out
O=gpurun_out/r2am_poly_sustained.txt
: > $O
for p in 0 4 2 1; do
  echo "== RNAMSM_COL_POLY=$p" >> $O
  RNAMSM_COL_POLY=$p timeout 300 python tools/col_sustain.py 512 256 2>&1 | grep -E "iters=(10|3000)" >> $O
  RNAMSM_COL_POLY=$p timeout 300 python tools/col_sustain.py 4096 128 2>&1 | grep -E "iters=(10|1000)" >> $O
done
cat $O
