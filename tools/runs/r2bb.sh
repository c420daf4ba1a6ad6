#!/bin/bash
# round-2 GPU call BB (8 GPUs): per-class time shares of the fused sharded forward (cfg5, cfg4)
mkdir -p gpurun_out
O=gpurun_out
for w in cfg5 cfg4; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29516 bench.py --gpus 8 --shard --fused --workload $w --steps 5 --warmup 3 --no-secondary --no-cpu-baseline > $O/r2bb_shard_$w.log 2>&1; echo "rc=$?" >> $O/r2bb_shard_$w.log
python - $w <<'PY'
import json,sys
for l in open(f"gpurun_out/r2bb_shard_{sys.argv[1]}.log"):
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']
        print(sys.argv[1], "ms", round(d['ms_per_step'],3), "e2e", d['e2e'].get('ms_per_step'))
        print("  shares", r['class_time_share']); print("  tflops", r['class_tflops'])
PY
done
