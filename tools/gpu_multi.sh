#!/usr/bin/env bash
# N-GPU visit: multi-process strip parity (frame-submit loop, gated and one-stream), bench at N.  usage: tools/gpu_multi.sh <tag> <N> [reps]
set -uo pipefail
OUT=gpurun_out/${1:-multi}; N=${2:-2}; REPS=${3:-1}; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 200 $TR --master-port 29551 tools/check_frame_strips.py 1280 720 6 2xf16 1 > $OUT/check_gated_n$N.log 2>&1; echo "strip frame-submit parity (gated, 720p) rc=$?"; grep "rank" $OUT/check_gated_n$N.log | tail -$N
timeout 200 $TR --master-port 29552 tools/check_frame_strips.py 640 400 6 f16 0 > $OUT/check_serial_n$N.log 2>&1; echo "strip frame-submit parity (one stream) rc=$?"; grep "rank" $OUT/check_serial_n$N.log | tail -$N
if [ "$N" = 2 ]; then timeout 500 python -m pytest tests -m gpu -q -x --timeout=300 -k "multi_process or two_gpus" > $OUT/pytest_multi.log 2>&1; echo "pytest multi rc=$?"; tail -3 $OUT/pytest_multi.log; fi
for rep in $(seq 1 $REPS); do
  timeout 300 $TR --master-port 2956$rep bench.py --gpus $N --steps 100 --warmup 5 > $OUT/bench_n${N}_$rep.json 2> $OUT/bench_n${N}_$rep.err; echo "bench N=$N rep $rep rc=$?"
done
timeout 300 $TR --master-port 29571 bench.py --gpus $N --steps 100 --warmup 5 --one-stream-strips > $OUT/bench_n${N}_onestream.json 2> $OUT/bench_n${N}_onestream.err; echo "bench one-stream rc=$?"
python - $OUT <<'PY'
import glob, json, os, sys
for p in sorted(glob.glob(os.path.join(sys.argv[1], "bench_n*.json"))):
    try:
        d = json.loads(open(p).read().strip().splitlines()[-1])
        print("%-28s %7.1f fps  e2e %7.1f (d2h %d)  convs %.3f ms  replicas %s" % (os.path.basename(p), d["value"], d["e2e"]["value"], d["e2e"]["d2h_bytes_per_step"], d["roofline"]["conv"]["ms"], d.get("replicas", {}).get("value")))
    except Exception as e:
        print(os.path.basename(p), "unreadable:", e, open(p[:-5] + ".err").read()[-400:])
PY
