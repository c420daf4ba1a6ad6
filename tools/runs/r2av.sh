#!/bin/bash
# round-2 GPU call AV (4 GPUs): the driver's N=4 bench line (final code)
mkdir -p gpurun_out
O=gpurun_out
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 4 --steps 10 --warmup 3 ) > $O/r2av_bench_n4.log 2>&1; echo "rc=$?" >> $O/r2av_bench_n4.log
python - <<'PY'
import json
for l in open("gpurun_out/r2av_bench_n4.log"):
    if l.startswith('{'):
        d=json.loads(l)
        print("N=4 ms", round(d['ms_per_step'],3), "tok/s", round(d['value']), "e2e", round(d['e2e']['value']))
        s=d['secondary']
        print({k:(v.get('sharded_ms'), v.get('speedup'), v.get('e2e_speedup')) for k,v in s.items() if k in ('cfg5','cfg4')}, s['parity'].get('parity_err'))
        print("cfg3", s['cfg3'].get('one_forward_per_msa'), s['cfg3'].get('forward_batch'))
PY
tail -4 $O/r2av_bench_n4.log | cut -c1-160
