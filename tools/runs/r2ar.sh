#!/bin/bash
# round-2 GPU call AR (1 GPU): compute-sanitizer over the new kernels (row_attn_short, register-cached row softmax)
mkdir -p gpurun_out
O=gpurun_out
SEL='row_attention_short and (7-36 or 130-64 or 61-100 or 300-128 or 1-20)'
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "$SEL or (row_attention_chain and (7-36 or 64-257))" > $O/r2ar_memcheck.log 2>&1; echo "memcheck rc=$?" >> $O/r2ar_memcheck.log
tail -6 $O/r2ar_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "row_attention_short and f16 and (7-36 or 61-100)" > $O/r2ar_racecheck.log 2>&1; echo "racecheck rc=$?" >> $O/r2ar_racecheck.log
tail -12 $O/r2ar_racecheck.log
timeout 600 compute-sanitizer --tool synccheck --error-exitcode 9 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "row_attention_short and f16 and (7-36 or 300-128)" > $O/r2ar_synccheck.log 2>&1; echo "synccheck rc=$?" >> $O/r2ar_synccheck.log
tail -5 $O/r2ar_synccheck.log
