#!/bin/bash
# round-2 GPU call AP (2 GPUs): sharded tests incl. the uneven shapes padded inside the sharded forward
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sharded.py -m gpu -q -x > gpurun_out/r2ap_sharded_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2ap_sharded_tests.log
tail -25 gpurun_out/r2ap_sharded_tests.log
