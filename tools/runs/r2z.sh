#!/bin/bash
# round-2 GPU call Z (1 GPU): col_attn_fa as the default: stand-alone numbers, whole GPU test suite, bench with secondary
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python tools/col_bench.py 512 256 4096 128 1024 1024 256 300 > $O/r2z_col_bench.txt 2>&1; cat $O/r2z_col_bench.txt
timeout 1200 python -m pytest tests -m gpu -x -q > $O/r2z_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2z_pytest.log
tail -4 $O/r2z_pytest.log
timeout 900 python bench.py --no-cpu-baseline > $O/r2z_bench.log 2>&1; echo "bench rc=$?" >> $O/r2z_bench.log
python - <<'PY'
import json
for l in open("gpurun_out/r2z_bench.log"):
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']
        print("cfg2 ms", round(d['ms_per_step'],3), "tok/s", round(d['value']), "e2e", round(d['e2e']['value']), d['clocks'])
        print(r['class_time_share']); print(r['class_tflops'])
        for k,v in d.get('secondary',{}).items():
            if isinstance(v,dict): print(k, {kk:vv for kk,vv in v.items() if kk in ('ms_per_step','tokens_per_s','class_tflops','whole_forward_tflops')})
PY
tail -2 $O/r2z_bench.log | cut -c1-300
