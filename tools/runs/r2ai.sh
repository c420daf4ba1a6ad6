#!/bin/bash
# round-2 GPU call AI (1 GPU): row softmax with the logits cached in registers; whole GPU suite; default bench with secondary
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > $O/r2ai_pytest.log 2>&1; echo "rc=$?" >> $O/r2ai_pytest.log
tail -4 $O/r2ai_pytest.log
timeout 900 python bench.py --no-cpu-baseline > $O/r2ai_bench.log 2>&1; echo "bench rc=$?" >> $O/r2ai_bench.log
python - <<'PY'
import json
for l in open("gpurun_out/r2ai_bench.log"):
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']
        print("cfg2 ms", round(d['ms_per_step'],3), "tok/s", round(d['value']), "e2e", round(d['e2e']['value']), d['clocks'])
        print(r['class_time_share']); print(r['class_tflops'])
        for k,v in d.get('secondary',{}).items():
            if isinstance(v,dict): print(k, {kk:vv for kk,vv in v.items() if kk in ('ms_per_step','tokens_per_s','class_time_share','whole_forward_tflops')})
PY
