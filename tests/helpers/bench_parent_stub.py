"""One torchrun worker of tests/test_bench_supervisor.py: calls bench.try_pipelined_strips like bench.py's main() does at N > 1 and reports."""
import importlib.util
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
spec = importlib.util.spec_from_file_location("ptd_bench", os.path.join(ROOT, "bench.py"))
bench = importlib.util.module_from_spec(spec)
spec.loader.exec_module(bench)
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
line = bench.try_pipelined_strips(["--gpus", str(world)], rank, world, child_timeout=90)
with open(os.path.join(sys.argv[1], "parent_%d.json" % rank), "w") as f:
    json.dump({"rank": rank, "line": line}, f)
