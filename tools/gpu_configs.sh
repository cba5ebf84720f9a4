#!/usr/bin/env bash
# side configs of BASELINE.json on N GPUs: usage tools/gpu_configs.sh <tag> <N> <config> [config...]
set -uo pipefail
OUT=gpurun_out/${1:-cfg}; N=${2:-1}; shift 2; mkdir -p $OUT
for cfg in "$@"; do
  if [ "$N" = 1 ]; then timeout 400 python bench.py --config $cfg --steps 60 --warmup 5 --no-cpu-baseline > $OUT/bench_${cfg}_n1.json 2> $OUT/bench_${cfg}_n1.err
  else timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29581 bench.py --gpus $N --config $cfg --steps 60 --warmup 5 > $OUT/bench_${cfg}_n$N.json 2> $OUT/bench_${cfg}_n$N.err; fi
  echo "$cfg N=$N rc=$?"
  python -c "
import json; d=json.loads(open('$OUT/bench_${cfg}_n$N.json').read().strip().splitlines()[-1]); print('$cfg', d['n_gpus'], 'value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), d['config']['workload'], d.get('modes'))"
done
