#!/usr/bin/env bash
# GPU-box experiment: rebuild pt_trace with different compile-time knobs and time the bench's path-trace kernels.
set -u
OUT=gpurun_out/${1:-sweep}; mkdir -p $OUT
for cfg in "-DTR_REFILL=22 -DTR_MIN_BLOCKS=1" "-DTR_REFILL=22 -DTR_MIN_BLOCKS=8" "-DTR_REFILL=28 -DTR_MIN_BLOCKS=1" "-DTR_REFILL=14 -DTR_MIN_BLOCKS=1" "-DTR_REFILL=28 -DTR_MIN_BLOCKS=8"; do
  touch ai_path_tracer_denoiser_b200/csrc/ptd_pt.cu
  make -C ai_path_tracer_denoiser_b200/csrc EXTRA="$cfg" > $OUT/build.log 2>&1 || { echo "build failed $cfg"; tail -5 $OUT/build.log; continue; }
  python bench.py --no-cpu-baseline --steps 15 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
k={x['kernel']:round(x['ms'],3) for x in d['roofline']['kernels']}
print('$cfg', round(d['value'],1), k)" | tee -a $OUT/sweep.log
done
