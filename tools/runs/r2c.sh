#!/bin/bash
# round-2 GPU call C (1 GPU): last-arriver fused residual+LayerNorm GEMM, 16-bit LM head
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -x -q -k "residual_layernorm or layernorm" > $O/r2c_lnfuse.log 2>&1; echo "rc=$?" >> $O/r2c_lnfuse.log
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2c_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2c_pytest.log
for fuse in 0 1; do
  RNAMSM_FUSE_LN=$fuse timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-secondary > $O/r2c_bench_f${fuse}.log 2>&1
done
tail -3 $O/r2c_lnfuse.log; tail -6 $O/r2c_pytest.log
for f in $O/r2c_bench_f*; do echo $f; python - "$f" <<'PY'
import json,sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']
        print(round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3), d['clocks']['sm_mhz'], r['class_time_share'], r['class_tflops'])
PY
done
