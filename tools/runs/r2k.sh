#!/bin/bash
# round-2 GPU call K (1 GPU): tf32x3 mode after the round-to-nearest split
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -q -s -k "tf32" > $O/r2k_tf32_ops.log 2>&1; echo "rc=$?" >> $O/r2k_tf32_ops.log
timeout 600 python -m pytest tests/test_gpu_model.py -q -s -k "tf32x3" > $O/r2k_tf32_model.log 2>&1; echo "rc=$?" >> $O/r2k_tf32_model.log
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2k_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2k_pytest.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-secondary --precision tf32x3 > $O/r2k_bench_tf32x3.log 2>&1
timeout 300 python __graft_entry__.py smoke > $O/r2k_smoke.log 2>&1
grep -E "tf32x3|passed|failed" $O/r2k_tf32_ops.log $O/r2k_tf32_model.log | grep -v "print" | head -20; tail -4 $O/r2k_pytest.log; tail -3 $O/r2k_smoke.log
python - <<'PY'
import json
for l in open("gpurun_out/r2k_bench_tf32x3.log"):
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']
        print(round(d['ms_per_step'],3), d['value'], r['class_time_share'], r['class_tflops'])
PY
