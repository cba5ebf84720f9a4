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


def test_autotune_switches_on_only_verified_faster_opt_ins(monkeypatch):
    """bench.run_autotune: an opt-in is enabled (with its best knob set) only if its self-check says bit-identical AND >= 3 % faster."""
    import importlib.util
    import types
    spec = importlib.util.spec_from_file_location("ptd_bench_autotune", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    for k in ("PTD_PT_RAY_SORT", "PTD_PT_RAY_SORT_REFILL", "PTD_PT_RAY_SORT_FROM", "PTD_PT_WIDE_LOOKBACK", "PTD_PT_SMEM_STACK", "PTD_DN_PDL"):
        monkeypatch.delenv(k, raising=False)
    calls = []

    def fake_run(cmd, **kw):
        feature = cmd[2]
        knobs = dict(a.split("=", 1) for a in cmd if "=" in a and a.startswith("PTD_"))
        calls.append((feature, knobs))
        table = {("ray_sort", ()): (True, 4.0, 3.5), ("ray_sort", (("PTD_PT_RAY_SORT_REFILL", "8"),)): (True, 4.0, 3.3),
                 ("ray_sort", (("PTD_PT_RAY_SORT_FROM", "1"),)): (True, 4.0, 3.6), ("wide_lookback", ()): (True, 4.0, 3.97), ("smem_stack", ()): (True, 4.0, 3.8), ("pdl", ()): (False, 0.8, 0.7)}
        if any(k in ("PTD_PT_SMEM_STACK", "PTD_PT_WIDE_LOOKBACK", "PTD_PT_RAY_SORT") for k in knobs):      # the combined check of the enabled set
            ok, base, feat = combined
        else:
            ok, base, feat = table[(feature, tuple(sorted(knobs.items())))]
        return types.SimpleNamespace(returncode=0, stdout="banner\n" + json.dumps({"feature": feature, "ok": ok, "base_ms": base, "feat_ms": feat}) + "\n", stderr="")

    monkeypatch.setattr(bench.subprocess, "run", fake_run)
    import argparse
    combined = (True, 4.0, 3.1)                                        # both together: better than the best single one (3.3)
    out = bench.run_autotune(argparse.Namespace(config="C3", mode="f16"))
    assert out["pt_combined"]["used"] and out["pt_combined"]["features"] == ["ray_sort", "smem_stack"]
    assert out["ray_sort"]["used"] and out["ray_sort"]["knobs"] == {"PTD_PT_RAY_SORT_REFILL": "8"} and len(out["ray_sort"]["tried"]) == 3
    assert os.environ.get("PTD_PT_RAY_SORT") == "1" and os.environ.get("PTD_PT_RAY_SORT_REFILL") == "8" and "PTD_PT_RAY_SORT_FROM" not in os.environ
    assert not out["wide_lookback"]["used"] and "PTD_PT_WIDE_LOOKBACK" not in os.environ          # bit-identical but only 1 % faster
    assert not out["pdl"]["used"] and "PTD_DN_PDL" not in os.environ                              # faster but not bit-identical
    assert out["smem_stack"]["used"] and os.environ.get("PTD_PT_SMEM_STACK") == "1"                 # bit-identical and 5 % faster
    for k in ("PTD_PT_RAY_SORT", "PTD_PT_RAY_SORT_REFILL", "PTD_PT_SMEM_STACK"):
        monkeypatch.delenv(k, raising=False)
    monkeypatch.setenv("PTD_DN_PDL", "0")                                                            # the caller decided: no self-check for it
    calls.clear()
    out = bench.run_autotune(argparse.Namespace(config="C3", mode="f16"))
    assert "pdl" not in [c[0] for c in calls] and out["pdl"]["used"] is False
    # ... and when the set is not better than its best member, only that member stays on
    for k in ("PTD_PT_RAY_SORT", "PTD_PT_RAY_SORT_REFILL", "PTD_PT_SMEM_STACK"):
        monkeypatch.delenv(k, raising=False)
    monkeypatch.delenv("PTD_DN_PDL", raising=False)
    combined = (True, 4.0, 3.4)
    out = bench.run_autotune(argparse.Namespace(config="C3", mode="f16"))
    assert not out["pt_combined"]["used"] and out["ray_sort"]["used"] and not out["smem_stack"]["used"]
    assert os.environ.get("PTD_PT_RAY_SORT") == "1" and "PTD_PT_SMEM_STACK" not in os.environ
    combined = (False, 4.0, 2.0)                                       # together they are not bit-identical: same fallback
    for k in ("PTD_PT_RAY_SORT", "PTD_PT_RAY_SORT_REFILL", "PTD_PT_SMEM_STACK"):
        monkeypatch.delenv(k, raising=False)
    out = bench.run_autotune(argparse.Namespace(config="C3", mode="f16"))
    assert not out["pt_combined"]["used"] and "PTD_PT_SMEM_STACK" not in os.environ and os.environ.get("PTD_PT_RAY_SORT") == "1"
    for k in ("PTD_PT_RAY_SORT", "PTD_PT_RAY_SORT_REFILL", "PTD_PT_SMEM_STACK"):
        os.environ.pop(k, None)
