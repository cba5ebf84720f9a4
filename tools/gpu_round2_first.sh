#!/usr/bin/env bash
# First GPU-box visit of round 2: validates, then times, the opt-ins written after round 1's GPU budget was spent.
# Every step runs under its own `timeout`; every device-side wait traps after 20 s, so nothing here can hang the box.
#   usage: gpurun --timeout 1500 -- 'bash tools/gpu_round2_first.sh r03a'           (1 GPU)
#          gpurun --gpus 4 --timeout 900 -- 'bash tools/gpu_round2_first.sh r03b strips 4'
set -uo pipefail
[ -x tools/microbench/umma_rate ] || nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/microbench/umma_rate tools/microbench/umma_rate.cu
TAG=${1:-r03a}; WHAT=${2:-single}; N=${3:-4}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > "$OUT/gpu.csv" 2>&1
B="--steps 100 --warmup 5 --no-cpu-baseline --no-autotune"      # explicit A/B runs below; the last single-GPU run is the plain default (autotune on)
if [ "$WHAT" = single ]; then
  timeout 120 tools/microbench/umma_rate 2048 > "$OUT/umma_rate.txt" 2>&1; echo "umma_rate rc=$?"; head -40 "$OUT/umma_rate.txt"
  PTD_OPTIN_TESTS=1 timeout 1200 python -m pytest tests -m gpu -q > "$OUT/pytest_gpu.log" 2>&1; echo "pytest rc=$?"; tail -4 "$OUT/pytest_gpu.log"
  timeout 400 python bench.py $B > "$OUT/bench_default.json" 2> "$OUT/bench_default.err"; echo "default rc=$?"
  for cfg in "4 22" "4 8" "5 22" "3 22"; do
    set -- $cfg
    PTD_PT_RAY_SORT=1 PTD_PT_RAY_SORT_BITS=$1 PTD_PT_RAY_SORT_REFILL=$2 timeout 400 python bench.py $B > "$OUT/bench_raysort_b$1_r$2.json" 2> "$OUT/bench_raysort_b$1_r$2.err"; echo "raysort bits=$1 refill=$2 rc=$?"
  done
  PTD_PT_RAY_SORT=1 PTD_PT_RAY_SORT_FROM=1 timeout 400 python bench.py $B > "$OUT/bench_raysort_from1.json" 2> "$OUT/bench_raysort_from1.err"
  PTD_PT_RAY_SORT=1 timeout 400 python bench.py $B --no-pipeline > "$OUT/bench_raysort_serial.json" 2> "$OUT/bench_raysort_serial.err"
  timeout 400 python bench.py $B --no-pipeline > "$OUT/bench_default_serial.json" 2> "$OUT/bench_default_serial.err"
  PTD_PT_SMEM_STACK=1 timeout 400 python bench.py $B > "$OUT/bench_smem_stack.json" 2> "$OUT/bench_smem_stack.err"; echo "smem stack rc=$?"
  PTD_PT_WIDE_LOOKBACK=1 timeout 400 python bench.py $B > "$OUT/bench_wide_lookback.json" 2> "$OUT/bench_wide_lookback.err"; echo "wide look-back rc=$?"
  PTD_DN_PDL=1 timeout 400 python bench.py $B > "$OUT/bench_pdl.json" 2> "$OUT/bench_pdl.err"; echo "pdl rc=$?"
  PTD_DN_PDL=1 timeout 400 python bench.py $B --no-pipeline > "$OUT/bench_pdl_serial.json" 2> "$OUT/bench_pdl_serial.err"
  for f in ray_sort wide_lookback smem_stack pdl; do timeout 200 python tools/selfcheck.py $f > "$OUT/selfcheck_$f.json" 2> "$OUT/selfcheck_$f.err"; echo "selfcheck $f rc=$?"; cat "$OUT/selfcheck_$f.json"; done
  timeout 900 python bench.py > "$OUT/bench_plain_default.json" 2> "$OUT/bench_plain_default.err"; echo "plain default (autotune + e2e auto) rc=$?"
  timeout 400 python bench.py $B --e2e fused > "$OUT/bench_e2e_fused.json" 2> "$OUT/bench_e2e_fused.err"; echo "fused rc=$?"
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
else
  TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
  timeout 300 $TR tools/check_strips_multi.py > "$OUT/check_n$N.log" 2>&1; echo "strip parity rc=$?"; tail -$N "$OUT/check_n$N.log"
  PTD_STRIP_PIPELINE=1 timeout 300 $TR tools/check_strips_multi.py 1280 720 4 > "$OUT/check_n${N}_gated_mail_720p.log" 2>&1; echo "strip parity with the gated mail rc=$?"; tail -$N "$OUT/check_n${N}_gated_mail_720p.log"
  timeout 900 $TR bench.py --gpus $N --steps 100 --warmup 5 > "$OUT/bench_n${N}_plain_default.json" 2> "$OUT/bench_n${N}_plain_default.err"; echo "plain default (supervised two-stream attempt) rc=$?"
  timeout 300 $TR bench.py --gpus $N $B > "$OUT/bench_n${N}_serial.json" 2> "$OUT/bench_n${N}_serial.err"; echo "serial rc=$?"
  for rep in 1 2 3; do
    PTD_STRIP_PIPELINE=1 timeout 300 $TR bench.py --gpus $N $B > "$OUT/bench_n${N}_pipelined_$rep.json" 2> "$OUT/bench_n${N}_pipelined_$rep.err"; echo "pipelined run $rep rc=$?"
  done
  PTD_DN_REPL_LEVEL=3 timeout 300 $TR bench.py --gpus $N $B > "$OUT/bench_n${N}_serial_repl.json" 2> "$OUT/bench_n${N}_serial_repl.err"; echo "replicated rc=$?"
  PTD_DN_REPL_LEVEL=3 PTD_STRIP_PIPELINE=1 timeout 300 $TR bench.py --gpus $N $B > "$OUT/bench_n${N}_pipelined_repl.json" 2> "$OUT/bench_n${N}_pipelined_repl.err"; echo "pipelined+replicated rc=$?"
  grep -ho '"value": [0-9.]*' "$OUT"/bench_n${N}_*.json | head -20
fi
ls -la "$OUT"
