#!/bin/bash
# round-2 GPU call AD (1 GPU): first run of row_attn_short (one-launch tied row attention for C <= 128)
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "row_attention" > $O/r2ad_row_tests.log 2>&1; echo "rc=$?" >> $O/r2ad_row_tests.log
tail -15 $O/r2ad_row_tests.log
timeout 600 python -m pytest tests/test_gpu_model.py -m gpu -q -x > $O/r2ad_model_tests.log 2>&1; echo "rc=$?" >> $O/r2ad_model_tests.log
tail -5 $O/r2ad_model_tests.log
for w in cfg1 cfg4; do
  timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-secondary --no-cpu-baseline > $O/r2ad_bench_$w.log 2>&1
  RNAMSM_ROW_SHORT=0 timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-secondary --no-cpu-baseline > $O/r2ad_bench_${w}_off.log 2>&1
done
python - <<'PY'
import json
for w in ("cfg1","cfg1_off","cfg4","cfg4_off"):
    for l in open(f"gpurun_out/r2ad_bench_{w}.log"):
        if l.startswith('{'):
            d=json.loads(l); r=d['roofline']
            print(w, "ms", round(d['ms_per_step'],3), "tok/s", round(d['value']), "e2e", round(d['e2e']['value']), d['clocks']['sm_mhz'])
            print("   ", r['class_time_share']); print("   ", r['class_tflops'])
PY
