#!/bin/bash
# round-2 GPU call F (2 GPUs): sharded single-MSA forward -- flag barriers, per-rank map rows to shared host memory
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi -L > $O/r2f_gpus.txt
timeout 900 python -m pytest tests/test_gpu_sharded.py -x -q > $O/r2f_pytest.log 2>&1; echo "rc=$?" >> $O/r2f_pytest.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29655"
timeout 900 $TR bench.py --gpus 2 --steps 10 --warmup 3 > $O/r2f_bench_n2.log 2>&1; echo "rc=$?" >> $O/r2f_bench_n2.log
timeout 600 $TR bench.py --gpus 2 --steps 5 --warmup 3 --shard --fused --workload cfg5 > $O/r2f_shard_cfg5_flag.log 2>&1
RNAMSM_NCCL_BARRIER=1 timeout 600 $TR bench.py --gpus 2 --steps 5 --warmup 3 --shard --fused --workload cfg5 > $O/r2f_shard_cfg5_nccl.log 2>&1
tail -4 $O/r2f_pytest.log
python - <<'PY'
import json
for f in ("r2f_bench_n2","r2f_shard_cfg5_flag","r2f_shard_cfg5_nccl"):
    for l in open(f"gpurun_out/{f}.log"):
        if l.startswith('{'):
            d=json.loads(l)
            print(f, round(d['ms_per_step'],3), 'e2e ms', round(d['e2e']['ms_per_step'],3), d['value'])
            if d.get('secondary'): print(json.dumps(d['secondary'])[:3000])
PY
tail -5 $O/r2f_bench_n2.log | cut -c1-400
