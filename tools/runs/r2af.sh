#!/bin/bash
# round-2 GPU call AF (1 GPU): row_attn_short with V prefetched behind the barriers; whole GPU suite; cfg1 / cfg4 bench
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > $O/r2af_pytest.log 2>&1; echo "rc=$?" >> $O/r2af_pytest.log
tail -5 $O/r2af_pytest.log
for w in cfg1 cfg4; do
  timeout 300 python bench.py --workload $w --steps 20 --warmup 5 --no-secondary --no-cpu-baseline > $O/r2af_bench_$w.log 2>&1
done
RNAMSM_ROW_SHORT=0 timeout 300 python bench.py --workload cfg1 --steps 20 --warmup 5 --no-secondary --no-cpu-baseline > $O/r2af_bench_cfg1_off.log 2>&1
python - <<'PY'
import json
for w in ("cfg1","cfg1_off","cfg4"):
    for l in open(f"gpurun_out/r2af_bench_{w}.log"):
        if l.startswith('{'):
            d=json.loads(l); r=d['roofline']
            print(w, "ms", round(d['ms_per_step'],3), "tok/s", round(d['value']), "e2e", round(d['e2e']['value']), d['clocks']['sm_mhz'])
            print("   ", r['class_time_share']); print("   ", r['class_tflops'])
PY
