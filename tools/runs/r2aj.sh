#!/bin/bash
# round-2 GPU call AJ (2 GPUs): sharded tests and the driver's N=2 bench line with the final kernels
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_sharded.py -m gpu -q -x > $O/r2aj_sharded_tests.log 2>&1; echo "rc=$?" >> $O/r2aj_sharded_tests.log
tail -4 $O/r2aj_sharded_tests.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > $O/r2aj_bench_n2.log 2>&1; echo "rc=$?" >> $O/r2aj_bench_n2.log
python - <<'PY'
import json
for l in open("gpurun_out/r2aj_bench_n2.log"):
    if l.startswith('{'):
        d=json.loads(l)
        print("N=2 ms", round(d['ms_per_step'],3), "tok/s", round(d['value']), "e2e", round(d['e2e']['value']))
        print(json.dumps(d.get('secondary'))[:1500])
PY
tail -3 $O/r2aj_bench_n2.log | cut -c1-300
