#!/bin/bash
# round-2 GPU call B (1 GPU): fused residual+LayerNorm GEMM, XU token ring in the column attention
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -x -q -k "residual_layernorm or layernorm" > $O/r2b_lnfuse.log 2>&1; echo "rc=$?" >> $O/r2b_lnfuse.log
for ring in 0 1; do
  RNAMSM_COL_RING=$ring timeout 300 python tools/col_bench.py > $O/r2b_colbench_ring$ring.txt 2>&1
done
RNAMSM_COL_RING=1 timeout 600 python -m pytest tests/test_gpu_ops.py -x -q -k "column_attention" > $O/r2b_colring_test.log 2>&1; echo "rc=$?" >> $O/r2b_colring_test.log
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2b_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2b_pytest.log
for fuse in 0 1; do for ring in 0 1; do
  RNAMSM_FUSE_LN=$fuse RNAMSM_COL_RING=$ring timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-secondary > $O/r2b_bench_f${fuse}_r${ring}.log 2>&1
done; done
tail -3 $O/r2b_lnfuse.log; cat $O/r2b_colbench_ring*.txt; tail -3 $O/r2b_colring_test.log; tail -4 $O/r2b_pytest.log
for f in $O/r2b_bench_f*; do echo $f; python - "$f" <<'PY'
import json,sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']
        print(round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3), d['clocks']['sm_mhz'], r['class_time_share'], r['class_tflops'])
PY
done
