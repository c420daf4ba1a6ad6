#!/bin/bash
# round-2 GPU call A (1 GPU): tests, micro-benchmarks, bench with the secondary block, reference arm
mkdir -p gpurun_out
df -h /dev/shm > gpurun_out/r2a_shm.txt 2>&1
nproc >> gpurun_out/r2a_shm.txt; free -g >> gpurun_out/r2a_shm.txt
timeout 120 tools/micro/mufu_bench > gpurun_out/r2a_mufu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2a_bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/r2a_bench.log
timeout 600 python bench.py --impl reference --steps 2 --warmup 0 > gpurun_out/r2a_ref.log 2>&1; echo "ref rc=$?" >> gpurun_out/r2a_ref.log
tail -3 gpurun_out/r2a_pytest.log; tail -c 1500 gpurun_out/r2a_bench.log
