"""bench.py end to end on a machine WITHOUT a GPU: the C ABI and torch.cuda are replaced by inert stand-ins, so that every line of the
benchmark's host logic - argument handling, the frame loop, the profiling leg, the e2e API self-checks and their fallbacks, the roofline
arithmetic, the JSON contract - executes here.  (What the numbers mean is the GPU's business; that this code runs is checked here.)"""
import contextlib
import importlib.util
import json
import os
import sys
import types

import numpy as np
import pytest
import torch

from conftest import ROOT


class _FakeLib:
    """The host-side entry points (scene loader, camera) are the real library's; every compute entry point (ptd_pt_*, ptd_dn_*, ptd_frame_*)
    returns PTD_OK without doing anything; `fail` names entry points that return PTD_ERR_CUDA instead."""
    def __init__(self, real, fail=()):
        self.real, self.fail, self.calls = real, set(fail), []

    def __getattr__(self, name):
        if not name.startswith(("ptd_pt_", "ptd_dn_", "ptd_frame_")) or name == "ptd_dn_strip_partition":     # pure host arithmetic
            return getattr(self.real, name)

        def f(*a):
            self.calls.append(name)
            return -4 if name in self.fail else 0
        return f


@pytest.fixture
def bench_env(monkeypatch):
    from ai_path_tracer_denoiser_b200 import capi, weights
    lib = _FakeLib(capi.lib())
    monkeypatch.setattr(capi, "lib", lambda: lib)
    monkeypatch.setattr(capi, "device_count", lambda: 1)

    class PT:
        def __init__(self, scene, device=0, flags=0, strip=None):
            cam = scene.camera[0]
            self.W, self.H = int(cam["res"][0]), int(cam["res"][1])
            self.P, self.depth, self.h = self.W * self.H, scene.counts()[3], 1
        def render(self, *a, **k): pass
        def profile(self, on=True): pass
        def launch_times(self): return np.full(2 * self.depth, 0.25, np.float32)
        def live_counts(self): return [self.P] + [self.P // 2] * (self.depth - 1), self.depth
        def launches(self): return 2 * self.depth

    class DN:
        def __init__(self, wfile, H, W, device=0, flags=1, strip=None):
            self.H, self.W, self.h = H, W, 2
        def padded_size(self): return (self.H + 31) // 32 * 32, (self.W + 31) // 32 * 32
        def forward(self, *a, **k): pass
        def profile(self, on=True): pass
        def launch_times(self): return [("pack_gbuffer", 0.01)] + [(n[0], 0.02) for n in weights.conv_layers()] + [("unpack_rgb", 0.01)]
        def launches(self): return 30

    monkeypatch.setattr(capi, "PathTracer", PT)
    monkeypatch.setattr(capi, "Denoiser", DN)

    class Ev:
        def __init__(self, enable_timing=False): pass
        def record(self, stream=None): pass
        def synchronize(self): pass
        def elapsed_time(self, other): return 100.0

    class St:
        cuda_stream = 0
        def __init__(self, priority=0): pass
        def wait_event(self, e): pass
        def synchronize(self): pass

    monkeypatch.setattr(torch.cuda, "get_device_properties", lambda d: types.SimpleNamespace(multi_processor_count=148))
    monkeypatch.setattr(torch.cuda, "set_device", lambda d: None)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a: None)
    monkeypatch.setattr(torch.cuda, "Event", Ev)
    monkeypatch.setattr(torch.cuda, "Stream", St)
    monkeypatch.setattr(torch.cuda, "stream", lambda s: contextlib.nullcontext())
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self: self)
    real_zeros = torch.zeros
    monkeypatch.setattr(torch, "zeros", lambda *a, **k: real_zeros(*a, **{kk: v for kk, v in k.items() if kk != "device"}))
    monkeypatch.setattr(torch, "empty", lambda *a, **k: real_zeros(*a, **{kk: v for kk, v in k.items() if kk != "device"}))   # inert entry points never fill the buffers
    spec = importlib.util.spec_from_file_location("ptd_bench_mock", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    monkeypatch.setattr(bench, "run_autotune", lambda args: {"ray_sort": {"used": False, "why": "mock"}})
    monkeypatch.setattr(bench.ClockSampler, "start", lambda self: None)
    monkeypatch.delenv("WORLD_SIZE", raising=False)
    return bench, lib


def _run(bench, capsys, argv):
    old = sys.argv
    sys.argv = ["bench.py"] + argv
    try:
        bench.main()
    finally:
        sys.argv = old
    lines = [ln for ln in capsys.readouterr().out.splitlines() if ln.startswith("{")]
    assert len(lines) == 1                                             # the contract: ONE JSON line on stdout
    return json.loads(lines[0])


def test_bench_host_logic_runs_and_keeps_the_json_contract(bench_env, capsys):
    bench, lib = bench_env
    d = _run(bench, capsys, ["--config", "C2", "--steps", "4", "--warmup", "3", "--no-cpu-baseline"])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
                "config", "gpu_launches", "e2e", "roofline", "clocks"):
        assert key in d, key
    assert d["metric"] == "denoised 720p frames/sec at 1spp (Sponza)" and d["n_gpus"] == 1 and d["steps"] == 4 and d["vs_baseline"] is None
    assert set(d["roofline"]) >= {"bound", "achieved", "peak", "unit", "frac", "traffic", "kernels"}
    assert set(d["e2e"]) >= {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step", "api", "self_check"}
    assert d["config"]["autotune"] == {"ray_sort": {"used": False, "why": "mock"}}
    # with inert entry points every API "reproduces" the two call sites, so auto picks the asynchronous one and balances submits and waits
    assert d["e2e"]["api"].startswith("ptd_frame_submit") and d["e2e"]["h2d_bytes_per_step"] == 84
    assert lib.calls.count("ptd_frame_submit") == lib.calls.count("ptd_frame_wait") == 3 + 3 + 4      # self-check + warm-up + timed steps


@pytest.mark.parametrize("broken,expect", [(("ptd_frame_submit",), "ptd_frame_host"), (("ptd_frame_wait",), "ptd_frame_host"),
                                           (("ptd_frame_submit", "ptd_frame_host"), "ptd_pt_render_host + ptd_dn_forward_host")])
def test_bench_e2e_falls_back_when_a_new_entry_point_fails(bench_env, capsys, broken, expect):
    bench, lib = bench_env
    lib.fail = set(broken)
    d = _run(bench, capsys, ["--config", "C2", "--steps", "2", "--warmup", "3", "--no-cpu-baseline"])
    assert d["e2e"]["api"].startswith(expect)
    assert any("not used" in v for v in d["e2e"]["self_check"].values())


def test_bench_explicit_e2e_modes(bench_env, capsys):
    bench, lib = bench_env
    for mode, api in (("calls", "ptd_pt_render_host"), ("fused", "ptd_frame_host"), ("async", "ptd_frame_submit")):
        d = _run(bench, capsys, ["--config", "C2", "--steps", "2", "--warmup", "3", "--no-cpu-baseline", "--e2e", mode, "--no-autotune"])
        assert d["e2e"]["api"].startswith(api) and "self_check" not in d["e2e"] and "autotune" not in d["config"]


@pytest.mark.parametrize("feature", ["ray_sort", "wide_lookback", "smem_stack", "pdl"])
def test_selfcheck_tool_runs(bench_env, capsys, feature, monkeypatch):
    """tools/selfcheck.py (what bench.py's autotune runs in subprocesses) under the same stand-ins: every feature's comparison and timing
    code executes and prints its one JSON line; the switch is set only while the second handle is created."""
    from ai_path_tracer_denoiser_b200 import capi
    seen = []
    PT, DN = capi.PathTracer, capi.Denoiser

    class PT2(PT):
        def __init__(self, *a, **k):
            super().__init__(*a, **k)
            seen.append(("pt", {k_: v for k_, v in os.environ.items() if k_.startswith("PTD_PT_") or k_ == "PTD_DN_PDL"}))
        def render_host(self, cam=None, iter=1): return np.zeros((10, self.H, self.W), np.float32)

    class DN2(DN):
        def __init__(self, *a, **k):
            super().__init__(*a, **k)
            seen.append(("dn", {k_: v for k_, v in os.environ.items() if k_.startswith("PTD_PT_") or k_ == "PTD_DN_PDL"}))
        def forward_host(self, g, reset): return np.zeros((3, self.H, self.W), np.float32)

    monkeypatch.setattr(capi, "PathTracer", PT2)
    monkeypatch.setattr(capi, "Denoiser", DN2)
    spec = importlib.util.spec_from_file_location("ptd_selfcheck_mock", os.path.join(ROOT, "tools", "selfcheck.py"))
    sc = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(sc)
    old = sys.argv
    sys.argv = ["selfcheck.py", feature, "--config", "C2", "--frames", "3", "--env", "PTD_PT_RAY_SORT_REFILL=8"]
    try:
        sc.main()
    finally:
        sys.argv = old
    d = json.loads([ln for ln in capsys.readouterr().out.splitlines() if ln.startswith("{")][-1])
    assert d["feature"] == feature and d["ok"] is True and d["base_ms"] > 0 and d["feat_ms"] > 0 and d["knobs"] == {"PTD_PT_RAY_SORT_REFILL": "8"}
    var = sc.SWITCH[feature]
    kind = "dn" if feature == "pdl" else "pt"
    handles = [env for k, env in seen if k == kind]
    assert var not in handles[-2] and handles[-1].get(var) == "1" and handles[-1].get("PTD_PT_RAY_SORT_REFILL") == "8"     # base, then the opt-in handle
    assert var not in os.environ and "PTD_PT_RAY_SORT_REFILL" not in os.environ


def test_bench_strip_mode_host_logic_runs(bench_env, capsys, monkeypatch):
    """The N > 1 leg of bench.py (one rank of a 2-rank job, torch.distributed replaced by no-ops, the supervised two-stream attempt switched
    off as the driver's child runs would have it): strips, serial frame loop, pipelined read-back, replicas context, JSON."""
    bench, lib = bench_env
    from ai_path_tracer_denoiser_b200 import capi, tiling
    import torch.distributed as dist
    PT, DN = capi.PathTracer, capi.Denoiser

    class PT2(PT):
        def __init__(self, scene, device=0, flags=0, strip=None):
            super().__init__(scene, device, flags, strip)
            if strip is not None:
                self.P = self.W * strip[1]
        def export_info(self): return b"p" * 8
        def connect(self, infos, rank): assert len(infos) == 2

    class DN2(DN):
        def export_info(self): return b"d" * 8
        def connect(self, infos, rank): assert len(infos) == 2

    monkeypatch.setattr(capi, "PathTracer", PT2)
    monkeypatch.setattr(capi, "Denoiser", DN2)
    monkeypatch.setattr(tiling, "exchange_blobs", lambda blob, d=None, world=1: [blob] * world)
    for name in ("init_process_group", "barrier", "destroy_process_group", "all_reduce"):
        monkeypatch.setattr(dist, name, lambda *a, **k: None)
    real_tensor = torch.tensor
    monkeypatch.setattr(torch, "tensor", lambda *a, **k: real_tensor(*a, **{kk: v for kk, v in k.items() if kk != "device"}))
    monkeypatch.setattr(torch, "device", lambda *a, **k: "cpu")
    for k, v in (("WORLD_SIZE", "2"), ("RANK", "0"), ("LOCAL_RANK", "0"), ("PTD_STRIP_PIPELINE", "0")):
        monkeypatch.setenv(k, v)
    d = _run(bench, capsys, ["--gpus", "2", "--config", "C2", "--steps", "3", "--warmup", "3"])
    assert d["n_gpus"] == 2 and d["scaling"] == "strong" and d["config"]["frame_loop"].startswith("serial")
    assert d["e2e"]["h2d_bytes_per_step"] == 84 and "replicas" in d and "autotune" not in d["config"] and "cpu_baseline" not in d
    monkeypatch.setenv("PTD_STRIP_PIPELINE", "1")                      # what the supervised child runs with
    d = _run(bench, capsys, ["--gpus", "2", "--config", "C2", "--steps", "3", "--warmup", "3"])
    assert d["config"]["frame_loop"].startswith("two streams") and "strip_loop" in d["config"]


def test_bench_mesh_config_adds_the_l1_model(bench_env, capsys):
    """C3 (a mesh scene): the roofline object also carries the host-side L1-wavefront model of pt_trace (ptd_bvh_probe_order is real host
    code even here)."""
    bench, lib = bench_env
    d = _run(bench, capsys, ["--config", "C3", "--steps", "2", "--warmup", "3", "--no-cpu-baseline", "--e2e", "calls", "--no-autotune"])
    m = d["roofline"]["l1_model"]
    assert "error" not in m and 60 < m["wavefronts_per_ray_binned"] < m["wavefronts_per_ray_arrival_order"] < 250 and m["frac"] > 0
    assert d["config"]["triangles"] > 200000
