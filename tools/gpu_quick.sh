#!/usr/bin/env bash
# quick 1-GPU visit: denoiser tests + conv timings per mode.  usage: tools/gpu_quick.sh <tag> [pytest -k expr]
set -uo pipefail
TAG=${1:-q}; KEXPR=${2:-}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
[ -x tools/microbench/umma_accum ] && timeout 100 tools/microbench/umma_accum > "$OUT/umma_accum.txt" 2>&1
if [ -n "$KEXPR" ]; then timeout 400 python -m pytest tests -m gpu -q -x --timeout=120 -k "$KEXPR" > "$OUT/pytest.log" 2>&1; else timeout 400 python -m pytest tests/test_gpu_dn.py tests/test_gpu_at_size.py -m gpu -q -s -x --timeout=120 > "$OUT/pytest.log" 2>&1; fi
echo "pytest rc=$?"; tail -5 "$OUT/pytest.log"; grep "max-abs / rel-L2" "$OUT/pytest.log"
for m in f16 tf32 2xf16 3xtf32; do
  timeout 150 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-side-modes --mode $m > "$OUT/bench_$m.json" 2> "$OUT/bench_$m.err"; echo "$m rc=$?"
done
python - "$OUT" <<'PY'
import glob, json, os, sys
for p in sorted(glob.glob(os.path.join(sys.argv[1], "bench_*.json"))):
    try:
        d = json.loads(open(p).read().strip().splitlines()[-1])
        print("%-24s %7.1f fps  e2e %7.1f  convs %.3f ms" % (os.path.basename(p), d["value"], d["e2e"]["value"], d["roofline"]["conv"]["ms"]))
        print("   ", {k: v for k, v in d["roofline"]["per_layer_ms"].items()})
    except Exception as e:
        print(os.path.basename(p), "unreadable:", e)
PY
