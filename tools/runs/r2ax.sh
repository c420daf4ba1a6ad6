#!/bin/bash
# round-2 GPU call AX (1 GPU): fc1 + GELU with 16 epilogue warps / 5-stage ring (kWide)
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "linear" > $O/r2ax_tests.log 2>&1; echo "rc=$?" >> $O/r2ax_tests.log; tail -3 $O/r2ax_tests.log
timeout 200 python tools/epi_cost_bench.py 131072 > $O/r2ax_epi_cost.txt 2>&1
timeout 200 python tools/epi_cost_bench.py 18432 >> $O/r2ax_epi_cost.txt 2>&1
echo "== RNAMSM_GELU_WIDE=0" >> $O/r2ax_epi_cost.txt
RNAMSM_GELU_WIDE=0 timeout 200 python tools/epi_cost_bench.py 131072 >> $O/r2ax_epi_cost.txt 2>&1
cat $O/r2ax_epi_cost.txt
for wv in 1 0; do
RNAMSM_GELU_WIDE=$wv timeout 600 python bench.py --steps 20 --warmup 5 --no-secondary --no-cpu-baseline > $O/r2ax_bench_$wv.log 2>&1
python - $wv <<'PY'
import json,sys
for l in open(f"gpurun_out/r2ax_bench_{sys.argv[1]}.log"):
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']
        print("WIDE="+sys.argv[1], "cfg2 ms", round(d['ms_per_step'],3), "tok/s", round(d['value']), "frac", r['frac'], r['class_tflops'])
PY
done
