#!/bin/bash
# round-2 GPU call BF (1 GPU): e2e path -- host-side token check, pitched DMA of the maps, final LayerNorm on MSA row 0 only
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_model.py -m gpu -q -x -k "streamed or cli or inference or extract" > $O/r2bf_tests.log 2>&1; echo "rc=$?" >> $O/r2bf_tests.log; tail -3 $O/r2bf_tests.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/r2bf_bench.log 2>&1
python - <<'PY'
import json
for l in open("gpurun_out/r2bf_bench.log"):
    if l.startswith('{'):
        d=json.loads(l)
        print("cfg2 ms", round(d['ms_per_step'],3), "tok/s", round(d['value']), "e2e", round(d['e2e']['value']), "e2e ms", round(d['e2e']['ms_per_step'],3), d['clocks']['sm_mhz'])
        for k,v in d['secondary'].items():
            if isinstance(v,dict) and 'ms_per_step' in v: print(k, v['ms_per_step'], v.get('e2e_ms_per_step'))
PY
