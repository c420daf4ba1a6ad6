#!/bin/bash
# round-2 GPU call I (8 GPUs): the driver's N=8 line with the secondary block; sharded cfg5 / cfg4 class shares
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi -L > $O/r2i_gpus.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29677"
timeout 900 $TR bench.py --gpus 8 --steps 10 --warmup 3 > $O/r2i_bench_n8.log 2>&1; echo "rc=$?" >> $O/r2i_bench_n8.log
timeout 600 $TR bench.py --gpus 8 --steps 5 --warmup 3 --shard --fused --workload cfg5 > $O/r2i_shard_cfg5_flag.log 2>&1
RNAMSM_NCCL_BARRIER=1 timeout 600 $TR bench.py --gpus 8 --steps 5 --warmup 3 --shard --fused --workload cfg5 > $O/r2i_shard_cfg5_nccl.log 2>&1
timeout 600 $TR bench.py --gpus 8 --steps 5 --warmup 3 --shard --fused --workload cfg4 > $O/r2i_shard_cfg4_flag.log 2>&1
python - <<'PY'
import json
for f in ("r2i_bench_n8","r2i_shard_cfg5_flag","r2i_shard_cfg5_nccl","r2i_shard_cfg4_flag"):
    for l in open(f"gpurun_out/{f}.log"):
        if l.startswith('{'):
            d=json.loads(l)
            print(f, round(d['ms_per_step'],3), 'e2e ms', round(d['e2e']['ms_per_step'],3), d['value'], d['roofline']['class_time_share'])
            if d.get('secondary'): print(json.dumps(d['secondary'])[:3500])
PY
tail -3 $O/r2i_bench_n8.log | cut -c1-300
