#!/bin/bash
# round-2 GPU call AY (1 GPU): where the time goes for cfg3 (64 alignments, depth 256, L 50-500)
mkdir -p gpurun_out
timeout 600 python bench.py --workload cfg3 --steps 2 --warmup 1 --no-secondary --no-cpu-baseline > gpurun_out/r2ay_cfg3.log 2>&1
python - <<'PY'
import json
for l in open("gpurun_out/r2ay_cfg3.log"):
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']
        print("cfg3 ms", round(d['ms_per_step'],2), "tok/s", round(d['value']), "e2e", round(d['e2e']['value']))
        print(r['class_time_share']); print(r['class_tflops'])
PY
timeout 200 python tools/col_bench.py 256 51 256 100 256 200 256 300 256 500 > gpurun_out/r2ay_col.txt 2>&1; cat gpurun_out/r2ay_col.txt
