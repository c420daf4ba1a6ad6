#!/bin/bash
# round-2 GPU call AT (1 GPU): what bounds the residual epilogue (K = 768: ~9 us per tile vs 3.3 us of MMAs)?
mkdir -p gpurun_out
O=gpurun_out/r2at_resid_epi.txt
: > $O
for d in 0 1 2 3 4 6; do
  RNAMSM_EPI_DEBUG=$d timeout 120 python tools/resid_epi_probe.py >> $O 2>&1
done
cat $O
