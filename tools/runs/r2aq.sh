#!/bin/bash
# round-2 GPU call AQ (1 GPU): does the fused residual + LayerNorm epilogue pay off when the stream is L2-resident (small T)?
mkdir -p gpurun_out
O=gpurun_out/r2aq_fuse_ln.txt
: > $O
for w in cfg1; do
  for f in 0 1; do
    RNAMSM_FUSE_LN=$f timeout 300 python bench.py --workload $w --steps 20 --warmup 5 --no-secondary --no-cpu-baseline > gpurun_out/r2aq_${w}_$f.log 2>&1
    python - "$w" "$f" >> $O <<'PY'
import json,sys
w,f=sys.argv[1],sys.argv[2]
for l in open(f"gpurun_out/r2aq_{w}_{f}.log"):
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']
        print(w, "FUSE_LN="+f, "ms", round(d['ms_per_step'],3), "tok/s", round(d['value']), r['class_time_share'], r['class_tflops'])
PY
  done
done
# other small shapes through the graph-free probe: eager forward at several (R, C)
for f in 0 1; do
  echo "== RNAMSM_FUSE_LN=$f" >> $O
  RNAMSM_FUSE_LN=$f timeout 300 python tools/graph_probe.py 256 64 256 100 128 300 256 200 512 128 2>&1 | grep "R=" >> $O
done
cat $O
