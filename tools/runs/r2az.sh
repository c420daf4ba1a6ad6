#!/bin/bash
# round-2 GPU call AZ (8 GPUs): sharded forward with the split count of the tied logits capped by the NVLink pull volume
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus 8 --steps 10 --warmup 3 > $O/r2az_bench_n8.log 2>&1; echo "rc=$?" >> $O/r2az_bench_n8.log
python - <<'PY'
import json
for l in open("gpurun_out/r2az_bench_n8.log"):
    if l.startswith('{'):
        d=json.loads(l)
        print("N=8 ms", round(d['ms_per_step'],3), "tok/s", round(d['value']), "e2e", round(d['e2e']['value']))
        s=d.get('secondary') or {}
        for k in ('cfg5','cfg4'):
            v=s.get(k,{})
            print(k, {kk:v.get(kk) for kk in ('single_gpu_ms','sharded_ms','sharded_e2e_ms','speedup','e2e_speedup','tokens_per_s','error')})
        print(s.get('parity',{}).get('parity_err'), s.get('cfg3',{}).get('one_forward_per_msa'))
PY
tail -2 $O/r2az_bench_n8.log | cut -c1-200
