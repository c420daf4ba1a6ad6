#!/bin/bash
# round-2 GPU call J (1 GPU): fp32 path on the tensor cores (tf32 x 3)
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -x -q -s -k "tf32" > $O/r2j_tf32_ops.log 2>&1; echo "rc=$?" >> $O/r2j_tf32_ops.log
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2j_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2j_pytest.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-secondary --precision fp32 > $O/r2j_bench_fp32_tensor.log 2>&1
RNAMSM_FP32_TENSOR=0 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-secondary --precision fp32 > $O/r2j_bench_fp32_ffma.log 2>&1
timeout 300 python __graft_entry__.py smoke > $O/r2j_smoke.log 2>&1
grep -E "tf32x3|passed|failed" $O/r2j_tf32_ops.log | head; tail -4 $O/r2j_pytest.log; tail -3 $O/r2j_smoke.log
for f in $O/r2j_bench_fp32_*; do echo $f; python - "$f" <<'PY'
import json,sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']
        print(round(d['ms_per_step'],3), d['value'], r['class_time_share'], r['class_tflops'])
PY
done
