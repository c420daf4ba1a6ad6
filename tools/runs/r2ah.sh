#!/bin/bash
# round-2 GPU call AH (1 GPU): CUDA-graph probe of the whole forward
mkdir -p gpurun_out
timeout 600 python tools/graph_probe.py > gpurun_out/r2ag_graph.txt 2>&1; tail -8 gpurun_out/r2ag_graph.txt
