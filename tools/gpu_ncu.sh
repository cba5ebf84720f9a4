#!/usr/bin/env bash
# ncu evidence of the round: launch list of the bench command + full captures of the dominant kernels (1 GPU).  usage: tools/gpu_ncu.sh <tag>
set -uo pipefail
OUT=gpurun_out/${1:-ncu}; mkdir -p $OUT
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-side-modes > $OUT/launches_bench.log 2>&1; echo "launches rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 28 -c 3 -f -o $OUT/conv_2xf16 python tools/run_frames.py C3 2 2xf16 > $OUT/ncu_conv.log 2>&1; echo "ncu conv rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:pt_trace -s 10 -c 1 -f -o $OUT/pt_trace python tools/run_frames.py C3 2 f16 > $OUT/ncu_pt.log 2>&1; echo "ncu pt_trace rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:pt_shade -s 10 -c 1 -f -o $OUT/pt_shade python tools/run_frames.py C3 2 f16 > $OUT/ncu_shade.log 2>&1; echo "ncu pt_shade rc=$?"
ls -la $OUT
