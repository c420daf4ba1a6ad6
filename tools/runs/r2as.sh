#!/bin/bash
# round-2 GPU call AS (1 GPU): whole GPU suite after the range guard / occupancy check / sharded padding changes
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2as_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2as_pytest.log
tail -6 gpurun_out/r2as_pytest.log
timeout 300 python -m pytest tests/test_gpu_model.py -m gpu -q -x -s -k "check_fp16_range" 2>&1 | grep "range guard"
