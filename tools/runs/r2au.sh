#!/bin/bash
# round-2 GPU call AU (1 GPU): ncu --set full of the rewritten row softmax (cfg2 and cfg5 shapes, inside a forward) and of
# row_attn_short at 4096 x 128
mkdir -p gpurun_out
O=gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:row_softmax_kernel' -s 4 -c 1 -o $O/r2au_prof_softmax_cfg2 python bench.py --steps 1 --warmup 1 --no-secondary --no-cpu-baseline > $O/r2au_ncu_a.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:row_softmax_kernel' -s 4 -c 1 -o $O/r2au_prof_softmax_cfg5 python bench.py --workload cfg5 --steps 1 --warmup 1 --no-secondary --no-cpu-baseline > $O/r2au_ncu_b.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:row_attn_short' -s 3 -c 1 -o $O/r2au_prof_row_short_cfg4 python tools/row_short_bench.py 4096 128 > $O/r2au_ncu_c.log 2>&1
ls -la $O | grep r2au
