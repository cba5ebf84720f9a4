#!/usr/bin/env bash
# GPU-box experiment: rebuild the path tracer with different compile-time knobs and time the bench's path-trace kernels.
set -u
OUT=gpurun_out/${1:-sweep}; mkdir -p $OUT
shift
for cfg in "$@"; do
  touch ai_path_tracer_denoiser_b200/csrc/ptd_pt.cu
  make -C ai_path_tracer_denoiser_b200/csrc EXTRA="$cfg" > $OUT/build.log 2>&1 || { echo "build failed $cfg"; tail -5 $OUT/build.log; continue; }
  python bench.py --no-cpu-baseline --steps 30 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
k={x['kernel']:round(x['ms'],3) for x in d['roofline']['kernels']}
print('$cfg', round(d['value'],1), k)" | tee -a $OUT/sweep.log
done
touch ai_path_tracer_denoiser_b200/csrc/ptd_pt.cu; make -C ai_path_tracer_denoiser_b200/csrc > /dev/null 2>&1
