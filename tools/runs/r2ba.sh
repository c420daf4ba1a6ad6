#!/bin/bash
# round-2 GPU call BA (1 GPU): barrier-counter ring of row_attn_short (two-stream test), row attention tests, model tests
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py -m gpu -q -x -k "row_attention or model or layer or batch" > gpurun_out/r2ba_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2ba_tests.log
tail -5 gpurun_out/r2ba_tests.log
