#!/bin/bash
# round-2 GPU call AN (1 GPU): the driver's commands as it runs them -- smoke, reference arm, default bench (CPU baseline + secondary incl. cfg3)
mkdir -p gpurun_out
O=gpurun_out
( time timeout 600 python -c "import __graft_entry__ as g; g.smoke()" ) > $O/r2an_smoke.log 2>&1; echo "smoke rc=$?" >> $O/r2an_smoke.log; tail -4 $O/r2an_smoke.log
( time timeout 900 python bench.py --impl reference --steps 3 --warmup 1 ) > $O/r2an_ref.log 2>&1; echo "ref rc=$?" >> $O/r2an_ref.log; tail -5 $O/r2an_ref.log | cut -c1-400
( time timeout 1200 python bench.py ) > $O/r2an_bench.log 2>&1; echo "bench rc=$?" >> $O/r2an_bench.log
python - <<'PY'
import json
for l in open("gpurun_out/r2an_bench.log"):
    if l.startswith('{'):
        d=json.loads(l)
        print("cfg2 ms", round(d['ms_per_step'],3), "tok/s", round(d['value']), "e2e", round(d['e2e']['value']), d['clocks'], "frac", d['roofline']['frac'])
        print("cpu", d['cpu_baseline'])
        print("cfg3", json.dumps(d['secondary'].get('cfg3')))
        print({k:(v.get('ms_per_step') if isinstance(v,dict) else v) for k,v in d['secondary'].items()})
PY
tail -5 $O/r2an_bench.log | cut -c1-200
