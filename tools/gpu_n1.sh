#!/usr/bin/env bash
# 1-GPU visit: the whole gpu suite (bounded), the default bench line, the reference arm.  usage: tools/gpu_n1.sh <tag>
set -uo pipefail
OUT=gpurun_out/${1:-n1}; mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -q -x --timeout=180 > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest.log
timeout 400 python bench.py --steps 100 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; tail -2 $OUT/bench.err; cut -c1-700 $OUT/bench.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 0 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "ref rc=$?"; cut -c1-300 $OUT/bench_ref.json
