#!/usr/bin/env bash
# lean N-GPU visit: strip parity of the frame-submit loop (gated mail, 720p, contract mode), then the bench line at N.
# usage: tools/gpu_final.sh <tag> <N> [extra SM split to compare]
set -uo pipefail
OUT=gpurun_out/${1:-final}; N=${2:-2}; CMP=${3:-}; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 150 $TR --master-port 29551 tools/check_frame_strips.py 1280 720 4 2xf16 1 > $OUT/check_gated_n$N.log 2>&1; echo "strip frame-submit parity (gated, 720p) rc=$?"; grep "rank" $OUT/check_gated_n$N.log | tail -$N
timeout 300 $TR --master-port 29561 bench.py --gpus $N --steps 100 --warmup 5 > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err; echo "bench N=$N rc=$?"
if [ -n "$CMP" ]; then PTD_FRAME_SM_SPLIT=$CMP timeout 300 $TR --master-port 29562 bench.py --gpus $N --steps 100 --warmup 5 --no-side-modes > $OUT/bench_n${N}_split$CMP.json 2> $OUT/bench_n${N}_split$CMP.err; echo "bench split $CMP rc=$?"; fi
python - $OUT <<'PY'
import glob, json, os, sys
for p in sorted(glob.glob(os.path.join(sys.argv[1], "bench_n*.json"))):
    try:
        d = json.loads(open(p).read().strip().splitlines()[-1])
        print("%-28s %7.1f fps  e2e %7.1f  replicas %s" % (os.path.basename(p), d["value"], d["e2e"]["value"], d.get("replicas", {}).get("value")))
    except Exception as e:
        print(os.path.basename(p), "unreadable:", e, open(p[:-5] + ".err").read()[-400:])
PY
