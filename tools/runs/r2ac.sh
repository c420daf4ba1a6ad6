#!/bin/bash
# round-2 GPU call AC (1 GPU): evidence for the final column-attention kernel -- launch lists of the default cfg2 bench
# and of cfg1, ncu --set full of col_attn_fa_kernel at the cfg2 and cfg4 shapes
mkdir -p gpurun_out
O=gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r2ac_launches_cfg2.csv \
  python bench.py --steps 2 --warmup 1 --no-secondary --no-cpu-baseline > $O/r2ac_b_cfg2.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r2ac_launches_cfg1.csv \
  python bench.py --workload cfg1 --steps 2 --warmup 1 --no-secondary --no-cpu-baseline > $O/r2ac_b_cfg1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:col_attn_fa_kernel' -s 3 -c 1 -o $O/r2ac_prof_col_fa_cfg2 python tools/col_bench.py 512 256 > $O/r2ac_ncu2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:col_attn_fa_kernel' -s 3 -c 1 -o $O/r2ac_prof_col_fa_cfg4 python tools/col_bench.py 4096 128 > $O/r2ac_ncu4.log 2>&1
timeout 300 python bench.py --workload cfg1 --steps 20 --warmup 5 --no-secondary --no-cpu-baseline > $O/r2ac_bench_cfg1.log 2>&1
tail -c 600 $O/r2ac_bench_cfg1.log
ls -la $O | grep r2ac
