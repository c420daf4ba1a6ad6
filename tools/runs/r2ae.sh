#!/bin/bash
# round-2 GPU call AE (1 GPU): row_attn_short with 8 epilogue warps / two-deep staging, batched split loads
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "row_attention" > $O/r2ae_row_tests.log 2>&1; echo "rc=$?" >> $O/r2ae_row_tests.log
tail -5 $O/r2ae_row_tests.log
for w in cfg1 cfg4; do
  timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-secondary --no-cpu-baseline > $O/r2ae_bench_$w.log 2>&1
done
python - <<'PY'
import json
for w in ("cfg1","cfg4"):
    for l in open(f"gpurun_out/r2ae_bench_{w}.log"):
        if l.startswith('{'):
            d=json.loads(l); r=d['roofline']
            print(w, "ms", round(d['ms_per_step'],3), "tok/s", round(d['value']), "e2e", round(d['e2e']['value']), d['clocks']['sm_mhz'])
            print("   ", r['class_time_share']); print("   ", r['class_tflops'])
PY
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:row_attn_short' -s 12 -c 1 -o $O/r2ae_prof_row_short_cfg1 python bench.py --workload cfg1 --steps 2 --warmup 1 --no-secondary --no-cpu-baseline > $O/r2ae_ncu.log 2>&1
