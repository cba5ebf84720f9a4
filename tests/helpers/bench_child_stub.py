"""Stand-in for the child benchmark of bench.py's try_pipelined_strips (tests/test_bench_supervisor.py): joins the CHILD process group that
the supervisor's environment describes (its own MASTER_PORT, no agent store), all-reduces over gloo, and rank 0 prints a JSON line.
argv[1]: "ok" | "fail1" (rank 1 exits 3 after the rendezvous) | "nojson" (rank 0 prints nothing)."""
import json
import os
import sys

import torch
import torch.distributed as dist

mode = sys.argv[1]
assert os.environ.get("PTD_STRIP_PIPELINE") == "1" and "TORCHELASTIC_USE_AGENT_STORE" not in os.environ
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
t = torch.tensor([rank + 1.0])
dist.all_reduce(t)
dist.destroy_process_group()
if mode == "fail1" and rank == 1:
    sys.exit(3)
if rank == 0 and mode != "nojson":
    print("some banner line")
    print(json.dumps({"metric": "stub", "value": float(t.item()), "port": os.environ["MASTER_PORT"]}))
