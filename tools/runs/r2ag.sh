#!/bin/bash
# round-2 GPU call AG (1 GPU): stand-alone timing of row_attn_short vs the chain; CUDA-graph probe of the whole forward
mkdir -p gpurun_out
timeout 300 python tools/row_short_bench.py > gpurun_out/r2ag_row_short.txt 2>&1; cat gpurun_out/r2ag_row_short.txt
timeout 600 python tools/graph_probe.py > gpurun_out/r2ag_graph.txt 2>&1; tail -8 gpurun_out/r2ag_graph.txt
