"""CPU test of bench.py's N > 1 supervisor (try_pipelined_strips): under a real torchrun launch with 2 workers, the opt-in two-stream strip
loop is attempted as a child process group on its own rendezvous port; rank 0 relays the child's JSON line only when EVERY rank's child
succeeded, otherwise every rank gets None and falls back to the serial loop in-process.  The child benchmark is replaced by a stub."""
import json
import os
import socket
import subprocess
import sys

import pytest

from conftest import ROOT

HELPERS = os.path.join(ROOT, "tests", "helpers")


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("mode,expect", [("ok", "relay"), ("fail1", "fallback"), ("nojson", "fallback")])
def test_supervisor_relays_or_falls_back(mode, expect, tmp_path):
    port = _free_port()
    env = dict(os.environ)
    env["PTD_BENCH_CHILD_CMD"] = "%s %s %s" % (sys.executable, os.path.join(HELPERS, "bench_child_stub.py"), mode)
    env.pop("PTD_STRIP_PIPELINE", None)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", str(port), os.path.join(HELPERS, "bench_parent_stub.py"), str(tmp_path)],
                       env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    res = [json.load(open(os.path.join(str(tmp_path), "parent_%d.json" % k))) for k in range(2)]
    if expect == "relay":
        line = json.loads(res[0]["line"])
        assert line["metric"] == "stub" and line["value"] == 3.0 and int(line["port"]) == port + 17     # the child group met on its own port
        assert res[1]["line"] == ""
    else:
        assert res[0]["line"] is None and res[1]["line"] is None
