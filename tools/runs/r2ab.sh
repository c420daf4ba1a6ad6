#!/bin/bash
# round-2 GPU call AB (1 GPU): is the in-forward column-attention rate (522 vs 745 TF/s alone) a sustained-clock effect?
mkdir -p gpurun_out
timeout 300 python tools/col_sustain.py 512 256 > gpurun_out/r2ab_sustain.txt 2>&1
timeout 300 python tools/col_sustain.py 4096 128 >> gpurun_out/r2ab_sustain.txt 2>&1
cat gpurun_out/r2ab_sustain.txt
