"""bench.py end to end on a machine WITHOUT a GPU: the C ABI's compute entry points and torch.cuda are replaced by inert stand-ins, so that
every line of the benchmark's host logic - argument handling, the ptd_frame_submit / ptd_frame_wait loops, the profiling leg, the side
modes, the roofline arithmetic, the JSON contract - executes here.  (What the numbers mean is the GPU's business; that this code runs is
checked here.)  The reference arm runs for real: it is CPU code."""
import importlib.util
import json
import os
import subprocess
import sys
import types

import numpy as np
import pytest
import torch

from conftest import ROOT


class _FakeLib:
    """The host-side entry points (scene loader, camera) are the real library's; every compute entry point returns PTD_OK without doing anything."""
    def __init__(self, real):
        self.real, self.calls = real, []

    def __getattr__(self, name):
        if not name.startswith(("ptd_pt_", "ptd_dn_", "ptd_frame_")) or name in ("ptd_dn_strip_partition", "ptd_frame_slots"):     # pure host arithmetic
            return getattr(self.real, name)

        def f(*a):
            self.calls.append(name)
            return 0
        return f


@pytest.fixture
def bench_env(monkeypatch):
    from ai_path_tracer_denoiser_b200 import capi, weights
    lib = _FakeLib(capi.lib())
    monkeypatch.setattr(capi, "lib", lambda: lib)
    monkeypatch.setattr(capi, "device_count", lambda: 1)
    log = []

    class PT:
        def __init__(self, scene, device=0, flags=0, strip=None):
            cam = scene.camera[0]
            self.W, self.H = int(cam["res"][0]), int(cam["res"][1])
            self.P, self.depth, self.h, self.flags = self.W * (strip[1] if strip else self.H), scene.counts()[3], 1, flags
            self.inflight = 0
        def frame_submit(self, dn, rgb_out, gbuf_out=None, cam=None, iter=1, reset=False):
            assert self.inflight < 3
            self.inflight += 1
            log.append(("submit", rgb_out is not None, gbuf_out is not None, reset))
        def frame_wait(self):
            assert self.inflight > 0
            self.inflight -= 1
            log.append(("wait",))
        def frame_timer_start(self): log.append(("t0",))
        def frame_timer_stop(self): return 100.0
        def profile(self, on=True): pass
        def launch_times(self): return np.full(2 * self.depth, 0.25, np.float32)
        def live_counts(self): return [self.P] + [self.P // 2] * (self.depth - 1), self.depth
        def launches(self): return 2 * self.depth
        def export_info(self): return b"p" * 8
        def connect(self, infos, rank): assert len(infos) == 2

    class DN:
        def __init__(self, wfile, H, W, device=0, flags=1, strip=None):
            self.H, self.W, self.h, self.flags = H, W, 2, flags
        def padded_size(self): return (self.H + 31) // 32 * 32, (self.W + 31) // 32 * 32
        def profile(self, on=True): pass
        def launch_times(self): return [("pack_gbuffer", 0.01)] + [(n[0], 0.02) for n in weights.conv_layers()] + [("unpack_rgb", 0.01)]
        def launches(self): return 30
        def export_info(self): return b"d" * 8
        def connect(self, infos, rank): assert len(infos) == 2

    monkeypatch.setattr(capi, "PathTracer", PT)
    monkeypatch.setattr(capi, "Denoiser", DN)
    monkeypatch.setattr(torch.cuda, "set_device", lambda d: None)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a: None)
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self: self)
    real_zeros = torch.zeros
    monkeypatch.setattr(torch, "zeros", lambda *a, **k: real_zeros(*a, **{kk: v for kk, v in k.items() if kk != "device"}))
    spec = importlib.util.spec_from_file_location("ptd_bench_mock", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    monkeypatch.setattr(bench.ClockSampler, "start", lambda self: None)
    for k in ("WORLD_SIZE", "RANK", "LOCAL_RANK"):
        monkeypatch.delenv(k, raising=False)
    return bench, lib, log


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
    bench, lib, log = bench_env
    d = _run(bench, capsys, ["--config", "C2", "--steps", "4", "--warmup", "3", "--no-cpu-baseline", "--side-configs", "C2"])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
                "config", "gpu_launches", "e2e", "roofline", "clocks", "modes"):
        assert key in d, key
    assert d["metric"] == "denoised 720p frames/sec at 1spp (Sponza)" and d["n_gpus"] == 1 and d["steps"] == 4 and d["vs_baseline"] is None
    assert d["dtype"].startswith("f32-equivalent") and set(d["modes"]) == {"f16", "tf32"}        # the contract mode is the benched one
    assert set(d["roofline"]) >= {"bound", "achieved", "peak", "unit", "frac", "traffic", "kernels", "conv"}
    assert set(d["e2e"]) >= {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step", "api"}
    assert d["e2e"]["api"].startswith("ptd_frame_submit") and d["e2e"]["h2d_bytes_per_step"] == 84 and d["e2e"]["d2h_bytes_per_step"] == 52 * 1280 * 720
    submits = [e for e in log if e[0] == "submit"]
    assert len(submits) == len([e for e in log if e[0] == "wait"])                               # every frame submitted is awaited
    assert len(submits) == (3 + 4) + (3 + 4) + 2 * (4 + 4) + (4 + 4)                              # value leg, e2e leg, two side modes, one side config
    assert isinstance(d["config"]["side_C2_fps"], float) and d["config"]["side_f16_fps"] == d["modes"]["f16"]["value"]
    assert sum(1 for e in submits if e[1] and e[2]) == 3 + 4                                      # only the e2e leg hands host buffers in
    assert d["gpu_launches"] == (2 * 8 + 30) * 4


def test_bench_strip_mode_host_logic_runs(bench_env, capsys, monkeypatch):
    """The N > 1 leg (one rank of a 2-rank job, torch.distributed replaced by no-ops): strips created with the gated live-count mail, the same
    frame-submit loop, replicas context, JSON."""
    bench, lib, log = bench_env
    from ai_path_tracer_denoiser_b200 import capi, tiling
    import torch.distributed as dist
    monkeypatch.setattr(tiling, "exchange_blobs", lambda blob, d=None, world=1: [blob] * world)
    for name in ("init_process_group", "barrier", "destroy_process_group", "all_reduce"):
        monkeypatch.setattr(dist, name, lambda *a, **k: None)
    real_tensor = torch.tensor
    monkeypatch.setattr(torch, "tensor", lambda *a, **k: real_tensor(*a, **{kk: v for kk, v in k.items() if kk != "device"}))
    monkeypatch.setattr(torch, "device", lambda *a, **k: "cpu")
    for k, v in (("WORLD_SIZE", "2"), ("RANK", "0"), ("LOCAL_RANK", "0")):
        monkeypatch.setenv(k, v)
    created = []
    PT = capi.PathTracer

    class PT2(PT):
        def __init__(self, scene, device=0, flags=0, strip=None):
            super().__init__(scene, device, flags, strip)
            created.append((flags, strip))
    monkeypatch.setattr(capi, "PathTracer", PT2)
    d = _run(bench, capsys, ["--gpus", "2", "--config", "C2", "--steps", "3", "--warmup", "3"])
    assert d["n_gpus"] == 2 and d["scaling"] == "strong" and d["config"]["frame_loop"].startswith("ptd_frame_submit") and "one stream" not in d["config"]["frame_loop"]
    assert created[0][0] == capi.PT_GATED_MAIL and created[0][1] is not None
    assert d["e2e"]["h2d_bytes_per_step"] == 84 and "replicas" in d and "modes" not in d and "cpu_baseline" not in d
    d = _run(bench, capsys, ["--gpus", "2", "--config", "C2", "--steps", "3", "--warmup", "3", "--one-stream-strips"])
    assert "one stream" in d["config"]["frame_loop"]


def test_reference_arm_runs_on_the_cpu_without_the_product_library():
    """`bench.py --impl reference` (the driver's reference arm) for real, on the CPU-runnable config: the reference's own CPU path trace
    (oracle/_ref) + the torch-CPU denoiser port, and nothing of libptd.so in the process."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "C2", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1 and len(r.stdout.strip().splitlines()) == 1                           # ONE line on stdout, the loader's chatter suppressed
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["value"] > 0 and d["e2e"]["value"] == d["value"] and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["cpu_baseline"]["kind"].startswith("reference") and d["cpu_baseline"]["cores"] >= 1 and "1280x720" in d["cpu_baseline"]["sample"]
    assert not any("libptd" in s for s in d["native_so_loaded"]) and any("libref_pt" in s for s in d["native_so_loaded"])
