#!/usr/bin/env bash
# One GPU-box visit: parity tests, the bench line, the ncu launch list of the same command and full ncu captures
# of the two dominant kernels.  Everything lands in gpurun_out/<tag>/.   usage: tools/gpu_round.sh <tag> [what...]
# what: tests bench launches ncu_conv ncu_pt smoke  (default: all)
set -uo pipefail
TAG=${1:-r01}; shift || true
WHAT=${*:-tests bench launches ncu_conv ncu_pt smoke}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > "$OUT/gpu.csv" 2>&1
for w in $WHAT; do
  case $w in
    tests)    timeout 1200 python -m pytest tests -m gpu -x -q > "$OUT/pytest_gpu.log" 2>&1; echo "pytest rc=$?" | tee -a "$OUT/pytest_gpu.log"; tail -5 "$OUT/pytest_gpu.log";;
    smoke)    timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > "$OUT/smoke.log" 2>&1; echo "smoke rc=$?"; tail -2 "$OUT/smoke.log";;
    bench)    timeout 900 python bench.py > "$OUT/bench.json" 2> "$OUT/bench.err"; echo "bench rc=$?"; cat "$OUT/bench.json"; tail -3 "$OUT/bench.err";;
    launches) timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file "$OUT/launches.csv" \
                python bench.py --steps 2 --warmup 3 --no-cpu-baseline > "$OUT/launches_bench.log" 2>&1; echo "launches rc=$?";;
    ncu_conv) timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_tc -s ${NCU_CONV_SKIP:-28} -c ${NCU_CONV_COUNT:-3} -f -o "$OUT/conv_tc" \
                python tools/run_frames.py C3 2 tf32 > "$OUT/ncu_conv.log" 2>&1; echo "ncu_conv rc=$?";;
    ncu_pt)   timeout 900 ncu --set full --clock-control none --import-source on -k regex:pt_trace -s ${NCU_PT_SKIP:-9} -c ${NCU_PT_COUNT:-2} -f -o "$OUT/pt_trace" \
                python tools/run_frames.py C3 2 tf32 > "$OUT/ncu_pt.log" 2>&1; echo "ncu_pt rc=$?";;
  esac
done
du -sh gpurun_out; ls -la "$OUT"
