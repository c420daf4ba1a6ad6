#!/bin/bash
# round-2 GPU call G (1 GPU): new parity tests, racecheck repro, ncu evidence at the cfg5 / cfg4 shapes, launch list
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2g_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2g_pytest.log
timeout 600 python -m pytest tests/test_gpu_model.py -q -s -k "contact_pairs or range_stress or beats_the_reference" > $O/r2g_newtests.log 2>&1
timeout 120 compute-sanitizer --tool racecheck tools/micro/tmem_alloc2_racecheck > $O/r2g_racecheck_repro.log 2>&1
# tied logits + AV at 1024 x 1024 (cfg5), fp16 is the second pass of gemm_bench: 13 launches per kernel per pass
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:umma_gemm_kernel<\(int\)[12]' -s 31 -c 1 -o $O/r2g_prof_tied_logits_cfg5 python tools/gemm_bench.py 1024 1024 > $O/r2g_ncu1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:umma_gemm_kernel<\(int\)2' -s 18 -c 1 -o $O/r2g_prof_tied_av_cfg5 python tools/gemm_bench.py 1024 1024 > $O/r2g_ncu2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:col_attn_ws_kernel' -s 5 -c 1 -o $O/r2g_prof_col_cfg4 python tools/col_bench.py 4096 128 > $O/r2g_ncu3.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 396 -c 270 --csv --log-file $O/r2g_launches.csv python bench.py --steps 2 --warmup 3 --no-secondary --no-cpu-baseline > $O/r2g_ncu4.log 2>&1
tail -4 $O/r2g_pytest.log; grep -E "^\[|passed|failed" $O/r2g_newtests.log | head -20; tail -8 $O/r2g_racecheck_repro.log; ls -la $O/r2g_prof*
