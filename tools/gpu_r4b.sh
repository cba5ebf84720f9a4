#!/usr/bin/env bash
# 1-GPU visit after the pipelined pt_shade: whole gpu suite, shade phase profiles (both kernels), frame spans, default bench + C2.
set -uo pipefail
OUT=gpurun_out/${1:-r4b}; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q -x --timeout=240 > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest.log
DBG=$PWD/ai_path_tracer_denoiser_b200/libptd_dbg.so
for t in 0 1; do PTD_LIBPTD=$DBG PTD_PT_SHADE_TILED=$t timeout 100 python tools/shade_prof.py 2>&1 | tail -1; done | tee $OUT/shade_phases.json
for m in f16 2xf16; do PTD_LIBPTD=$DBG timeout 100 python tools/frame_spans.py $m 2>&1 | tail -1; done | tee $OUT/frame_spans.json
timeout 400 python bench.py --steps 100 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; tail -2 $OUT/bench.err; cut -c1-400 $OUT/bench.json
timeout 200 python bench.py --config C2 --steps 100 --warmup 5 --no-cpu-baseline --no-side-modes > $OUT/bench_c2.json 2> $OUT/bench_c2.err; echo "c2 rc=$?"; cut -c1-300 $OUT/bench_c2.json
PTD_PT_SHADE_TILED=1 timeout 200 python bench.py --config C2 --steps 100 --warmup 5 --no-cpu-baseline --no-side-modes > $OUT/bench_c2_tiled.json 2> $OUT/bench_c2_tiled.err; echo "c2 tiled rc=$?"; cut -c1-300 $OUT/bench_c2_tiled.json
