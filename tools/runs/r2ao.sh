#!/bin/bash
# round-2 GPU call AO (2 GPUs): N=2 bench line with the cfg3 farm in secondary
mkdir -p gpurun_out
O=gpurun_out
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 10 --warmup 3 ) > $O/r2ao_bench_n2.log 2>&1; echo "rc=$?" >> $O/r2ao_bench_n2.log
python - <<'PY'
import json
for l in open("gpurun_out/r2ao_bench_n2.log"):
    if l.startswith('{'):
        d=json.loads(l)
        print("N=2 ms", round(d['ms_per_step'],3), "tok/s", round(d['value']), "e2e", round(d['e2e']['value']))
        s=d['secondary']
        print("cfg3", json.dumps(s.get('cfg3'))[:700])
        print({k:(v.get('speedup'), v.get('e2e_speedup')) for k,v in s.items() if k in ('cfg5','cfg4')}, s['parity'].get('parity_err'))
PY
tail -6 $O/r2ao_bench_n2.log | cut -c1-200
