#!/usr/bin/env bash
# timing experiments with PTD_DN_DEBUG (garbage results): per-layer times of level 0 under each knob
OUT=gpurun_out/${1:-dbg}; mkdir -p $OUT
for dbg in 0 1 2 4 3 7; do
  for m in f16 tf32; do
    PTD_DN_DEBUG=$dbg timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-side-modes --mode $m > $OUT/b_${m}_$dbg.json 2> $OUT/b_${m}_$dbg.err
    python - $OUT/b_${m}_$dbg.json $m $dbg <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); pl=d["roofline"]["per_layer_ms"]
    print(sys.argv[2], "dbg", sys.argv[3], "convs %.3f" % d["roofline"]["conv"]["ms"], {k:pl[k] for k in ["enc1.l1","enc1.l2a","enc1.l2b","enc2.l2a","dec1.c1","dec1.c2","enc4.l2a","bott.l2a"]})
except Exception as e: print(sys.argv[2], sys.argv[3], "failed", e)
PY
  done
done
