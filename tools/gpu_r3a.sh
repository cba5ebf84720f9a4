#!/usr/bin/env bash
# Round-2 first GPU visit (1 GPU): microbench, the WHOLE gpu test suite with every opt-in un-skipped + the at-size parity tests,
# then A/B timings of every opt-in on the bench workload.  Each step under its own timeout.
set -uo pipefail
TAG=${1:-r3a}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > "$OUT/gpu.csv" 2>&1
nproc > "$OUT/nproc.txt"
timeout 120 tools/microbench/umma_rate 2048 > "$OUT/umma_rate.txt" 2>&1; echo "umma_rate rc=$?"; head -60 "$OUT/umma_rate.txt"
PTD_OPTIN_TESTS=1 timeout 1500 python -m pytest tests -m gpu -q -s --durations=15 > "$OUT/pytest_gpu.log" 2>&1; echo "pytest rc=$?"; tail -40 "$OUT/pytest_gpu.log"
B="--steps 60 --warmup 5 --no-cpu-baseline"
run() { name=$1; shift; timeout 300 env "$@" python bench.py $B $EXTRA > "$OUT/bench_$name.json" 2> "$OUT/bench_$name.err"; echo "$name rc=$?"; }
EXTRA=""
run default X=1
run raysort_b4 PTD_PT_RAY_SORT=1 PTD_PT_RAY_SORT_BITS=4
run raysort_b5 PTD_PT_RAY_SORT=1 PTD_PT_RAY_SORT_BITS=5
run raysort_b3 PTD_PT_RAY_SORT=1 PTD_PT_RAY_SORT_BITS=3
run raysort_from1 PTD_PT_RAY_SORT=1 PTD_PT_RAY_SORT_FROM=1
run smem_stack PTD_PT_SMEM_STACK=1
run smem_raysort PTD_PT_SMEM_STACK=1 PTD_PT_RAY_SORT=1
run wide_lookback PTD_PT_WIDE_LOOKBACK=1
run pdl PTD_DN_PDL=1
EXTRA="--mode tf32"
run tf32 X=1
EXTRA="--mode 3xtf32"
run 3xtf32 X=1
EXTRA="--no-pipeline"
run serial X=1
run serial_pdl PTD_DN_PDL=1
EXTRA="--e2e fused"
run e2e_fused X=1
EXTRA="--e2e async"
run e2e_async X=1
python - "$OUT" <<'PY'
import glob, json, os, sys
for p in sorted(glob.glob(os.path.join(sys.argv[1], "bench_*.json"))):
    try:
        d = json.loads(open(p).read().strip().splitlines()[-1])
        print("%-34s %7.1f fps  e2e %7.1f  convs %.3f ms  trace/bounce %s  shade/bounce %s" % (os.path.basename(p), d["value"], d["e2e"]["value"], d["roofline"]["conv"]["ms"],
              d["roofline"]["per_bounce_ms"]["pt_trace"], d["roofline"]["per_bounce_ms"]["pt_shade"]))
    except Exception as e:
        print(os.path.basename(p), "unreadable:", e)
PY
ls -la "$OUT"
