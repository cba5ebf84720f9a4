#!/usr/bin/env bash
# 1-GPU visit for pt_shade work: PT parity tests, then pipelined vs tiled shade timings.  usage: tools/gpu_shade.sh <tag> [pytest -k expr]
set -uo pipefail
TAG=${1:-shade}; KEXPR=${2:-}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
timeout 500 python -m pytest tests/test_gpu_pt.py -m gpu -q -x --timeout=100 ${KEXPR:+-k "$KEXPR"} > "$OUT/pytest.log" 2>&1
echo "pytest rc=$?"; tail -15 "$OUT/pytest.log"
for v in 0 1; do
  PTD_PT_SHADE_TILED=$v timeout 200 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-side-modes --mode f16 > "$OUT/bench_tiled$v.json" 2> "$OUT/bench_tiled$v.err"; echo "tiled=$v rc=$?"
  python - "$OUT/bench_tiled$v.json" <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); t = d["roofline"]["per_bounce_ms"]
print(d["value"], "shade", t["pt_shade"], round(sum(t["pt_shade"]), 4), "trace", round(sum(t["pt_trace"]), 4))
PY
done
