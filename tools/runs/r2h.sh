#!/bin/bash
# round-2 GPU call H (1 GPU): programmatic dependent launch A/B, tests, ncu evidence at cfg5 / cfg4 shapes
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2h_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2h_pytest.log
for pdl in 0 1; do
  RNAMSM_PDL=$pdl timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-secondary > $O/r2h_bench_pdl${pdl}.log 2>&1
  RNAMSM_PDL=$pdl timeout 600 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-secondary --workload cfg1 > $O/r2h_bench_cfg1_pdl${pdl}.log 2>&1
done
tools/micro/tmem_alloc2_racecheck > $O/r2h_repro_plain.log 2>&1; echo "rc=$?" >> $O/r2h_repro_plain.log
timeout 120 compute-sanitizer --tool racecheck --print-limit 20 tools/micro/tmem_alloc2_racecheck > $O/r2h_racecheck_repro.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:umma_gemm_kernel -s 135 -c 1 -o $O/r2h_prof_tied_logits_cfg5 python tools/gemm_bench.py 1024 1024 > $O/r2h_ncu1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:umma_gemm_kernel -s 148 -c 1 -o $O/r2h_prof_tied_av_cfg5 python tools/gemm_bench.py 1024 1024 > $O/r2h_ncu2.log 2>&1
tail -4 $O/r2h_pytest.log; cat $O/r2h_repro_plain.log; tail -5 $O/r2h_racecheck_repro.log
for f in $O/r2h_bench_*; do echo $f; python - "$f" <<'PY'
import json,sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']
        print(round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3), d['clocks']['sm_mhz'], d['value'])
PY
done
ls -la $O/r2h_prof*
