#!/bin/bash
# round-2 GPU call BE (1 GPU): full-size check of the one-launch tied row attention
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_zz_fullsize_gpu.py -m gpu -q -x -k "one_launch or tied_row" > gpurun_out/r2be_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2be_tests.log
tail -6 gpurun_out/r2be_tests.log
