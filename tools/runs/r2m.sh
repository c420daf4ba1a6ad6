#!/bin/bash
# round-2 GPU call M (1 GPU): softmax-step microbenchmark + the two-item-group column attention (RNAMSM_COL_GROUPS=2)
mkdir -p gpurun_out
O=gpurun_out
timeout 120 tools/micro/tmem_umma_bench > $O/r2m_tmem_umma.txt 2>&1; echo "rc=$?" >> $O/r2m_tmem_umma.txt
tail -8 $O/r2m_tmem_umma.txt
RNAMSM_COL_GROUPS=2 timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q -k "col" > $O/r2m_col_tests_g2.log 2>&1; echo "rc=$?" >> $O/r2m_col_tests_g2.log
tail -3 $O/r2m_col_tests_g2.log
timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q -k "col" > $O/r2m_col_tests_g1.log 2>&1; echo "rc=$?" >> $O/r2m_col_tests_g1.log
tail -3 $O/r2m_col_tests_g1.log
SH="512 256 4096 128 1024 1024 256 300 384 200 768 64"
echo "groups=1" > $O/r2m_col_bench.txt; timeout 300 python tools/col_bench.py $SH >> $O/r2m_col_bench.txt 2>&1
echo "groups=2" >> $O/r2m_col_bench.txt; RNAMSM_COL_GROUPS=2 timeout 300 python tools/col_bench.py $SH >> $O/r2m_col_bench.txt 2>&1
echo "groups=1 NT=2 forced" >> $O/r2m_col_bench.txt; RNAMSM_COL_NT=2 timeout 300 python tools/col_bench.py $SH >> $O/r2m_col_bench.txt 2>&1
cat $O/r2m_col_bench.txt
