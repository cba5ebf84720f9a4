#!/usr/bin/env python
"""bench.py - denoised 720p frames/s at 1 spp on the (procedural) Sponza-like scene, B200.

One "step" = one frame of the reference's render loop (main.cpp:120-168): 1-spp path trace (HP-1) of the camera of
frame k of a 300-frame pan, then the recurrent denoiser forward (HP-2) with the hidden state carried from frame k-1.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--mode 2xf16|f16|tf32|3xtf32|fp32] [--config C2..C5]

Prints ONE JSON line (see the contract in DESIGN.md / the task statement):
  value      frames/s, G-buffer and frames resident in HBM (kernel time only, CUDA events on the launch stream)
  e2e        the same frames through the host-pointer C ABI the reference's loop binds (ptd_pt_render_host +
             ptd_dn_forward_host: D2H of the 40*P-byte G-buffer, H2D of it again, D2H of the 12*P-byte frame - what
             pathtrace.cu:525 and main.cpp:104-105,91 do)
  roofline   dominant kernel (by share of the step) against its bound; `kernels` lists every kernel class
  cpu_baseline  the reference's own CPU path (oracle/_ref brute-force path trace + the oracle's torch-CPU denoiser) on a
             bounded sample, timed on this box's host cores
`--impl reference` times only that CPU path (rank 0 only) and prints the same line shape with "impl": "reference".
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "denoised 720p frames/sec at 1spp (Sponza)"
CONV_FLOP_PER_PX = None  # filled from the layer table


def layer_table(Hp, Wp):
    """(name, flops, bytes) per conv, SURVEY.md section 8a/8d: 2*9*Cin*Cout*H*W unpadded; (Cin+Cout)*H*W*4 bytes."""
    from ai_path_tracer_denoiser_b200.weights import conv_layers
    out = {}
    for name, _, _, ci, co, _ in conv_layers():
        if name.startswith("enc"):
            lvl = int(name[3]) - 1
        elif name.startswith("bott"):
            lvl = 5
        else:
            lvl = int(name[3]) - 1
        px = (Hp >> lvl) * (Wp >> lvl)
        out[name] = (2.0 * 9 * ci * co * px, 4.0 * (ci + co) * px)
    return out


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return dict(hbm=float(d["hbm_gbs"]), bf16=float(d["bf16_tflops"]), bf16_sustained=float(d.get("bf16_tflops_sustained", d["bf16_tflops"])),
                        source="MEASURED_PEAKS.json")
        except Exception:
            pass
    # the driver-written file is git-ignored and absent here; these are its values as recorded in BASELINE.md section 2
    return dict(hbm=6539.2, bf16=1633.5, bf16_sustained=1353.6, source="BASELINE.md section 2 (copy of MEASURED_PEAKS.json of this pool)")


class ClockSampler:
    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index),
                                          "--query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
                                          "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=3)
            except Exception:
                self.proc.kill()
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        reasons = []
        for i, nm in ((3, "hw_slowdown"), (4, "hw_thermal_slowdown"), (5, "sw_thermal_slowdown"), (6, "sw_power_cap")):
            if any(len(r) >= 7 and r[i].lower().startswith("active") for r in self.rows):
                reasons.append(nm)
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None, reasons=reasons, samples=len(sm))


def make_scene(config, W, H):
    from ai_path_tracer_denoiser_b200 import scenegen
    d = os.path.join(tempfile.gettempdir(), "ptd_bench_scenes_r%s" % os.environ.get("RANK", "0"))     # one directory per rank: no file races under torchrun
    os.makedirs(d, exist_ok=True)
    path, desc = scenegen.make_config(d, config)
    return path, desc


def cpu_reference_frame(scene_path, cam, live_counts, H, W, threads, row_stride, dn_runs=1):
    """Reference CPU path for one frame, bounded sample.  Path trace: oracle/_ref (the reference's own __host__ __device__
    intersection code, brute force over every face like pathtrace.cu:258-269, OpenMP) on the first bounce of every
    `row_stride`-th image row, scaled by (sum of live paths over the bounces) / (rays in the sample).  Denoise: the oracle's
    torch-CPU forward (oracle/dn_oracle.py, pinned against the reference model) on the full padded frame."""
    from oracle import reflib
    from oracle.dn_oracle import DenoiserOracle, synthetic_gbuffer
    from ai_path_tracer_denoiser_b200 import weights
    os.environ.setdefault("OMP_NUM_THREADS", str(threads))
    R = reflib.RefLib()
    devnull, saved = os.open(os.devnull, os.O_WRONLY), os.dup(1)
    os.dup2(devnull, 1)
    try:
        s = R.load_scene(scene_path)
    finally:
        os.dup2(saved, 1)
    R.set_camera(s, cam)
    rays, ms = R.cpu_first_bounce_rows(s, row_stride)
    pt_s = ms * 1e-3 * (float(sum(live_counts)) / rays)
    import torch
    torch.set_num_threads(threads)
    O = DenoiserOracle(weights.synthetic_state_dict(1234), threads=threads)
    x = synthetic_gbuffer(H, W, seed=1)
    O.forward(x, reset=True)
    t0 = time.perf_counter()
    for _ in range(dn_runs):
        O.forward(x, reset=False)
    dn_s = (time.perf_counter() - t0) / dn_runs
    return pt_s, dn_s, rays


def run_reference(args, W, H, config):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from ai_path_tracer_denoiser_b200 import capi
    threads = os.cpu_count() or 1
    scene_path, desc = make_scene(config, W, H)
    sc = capi.Scene(path=scene_path)
    nfaces = sc.counts()[2]
    # live-path profile of this scene/camera: from the survey's probe ratio when no GPU run is at hand (2.66 P at 720p Cornell;
    # measured 6.6 P for the enclosed hall) - the reference arm must not touch our kernels, so it uses the conservative P * depth bound / 2
    P = W * H
    live = [P] + [int(P * 0.8)] * 7 if nfaces else [P, int(.46 * P), int(.32 * P), int(.25 * P), int(.2 * P), int(.17 * P), int(.14 * P), int(.11 * P)]
    row_stride = max(1, H // 64) if nfaces else max(1, H // 256)
    times = []
    for k in range(args.warmup + args.steps):
        cam = capi.frame_camera(sc.camera[0], k)
        pt_s, dn_s, rays = cpu_reference_frame(scene_path, cam, live, H, W, threads, row_stride)
        if k >= args.warmup:
            times.append(pt_s + dn_s)
    t = float(np.mean(times))
    fps = 1.0 / t
    sample = "per step: brute-force first-bounce intersect of every %d-th row (%d rays x %d faces) scaled to sum(live paths)=%.2f*P, + one full %dx%d torch-CPU forward" % (
        row_stride, rays, nfaces, sum(live) / P, (H + 31) // 32 * 32, (W + 31) // 32 * 32)
    print(json.dumps({"metric": METRIC, "value": fps, "unit": "frames/s", "impl": "reference", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                      "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                      "config": {"workload": "%s: %s" % (config, desc), "note": "reference CPU path (oracle/_ref + oracle/dn_oracle.py), host cores only"},
                      "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": threads, "kind": "reference(path trace)+port(denoiser)", "sample": sample},
                      "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def run_autotune(args):
    """Opt-in code paths that were written after round 1's GPU budget was spent are switched on for this run only if,
    on THIS box and on the bench workload, tools/selfcheck.py finds them bit-identical to the default path and at least 3 % faster.
    Each check runs in its own process under a timeout, before this process touches the GPU: a fault or a hang in an opt-in costs its
    gain, never the benchmark.  Returns {feature: outcome} for the JSON line."""
    out = {}
    features = (("ray_sort", "PTD_PT_RAY_SORT", [{}, {"PTD_PT_RAY_SORT_REFILL": "8"}, {"PTD_PT_RAY_SORT_FROM": "1"}]),     # knob variants tried once the plain one passed
                ("wide_lookback", "PTD_PT_WIDE_LOOKBACK", [{}]), ("smem_stack", "PTD_PT_SMEM_STACK", [{}]), ("pdl", "PTD_DN_PDL", [{}]))
    t_begin = time.time()
    for feature, var, variants in features:
        if time.time() - t_begin > 240:                                  # the whole bench has to finish within minutes
            out[feature] = {"used": False, "why": "autotune time budget (240 s) spent"}
            continue
        if var in os.environ:                                            # the caller decided
            out[feature] = {"used": os.environ[var] not in ("", "0"), "why": "%s set by the caller" % var}
            continue
        best, tried = None, []
        for knobs in variants:
            try:
                cmd = [sys.executable, os.path.join(ROOT, "tools", "selfcheck.py"), feature, "--config", args.config, "--mode", args.mode]
                for k, v in knobs.items():
                    cmd += ["--env", "%s=%s" % (k, v)]
                r = subprocess.run(cmd, capture_output=True, text=True, timeout=150)
                line = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
                if r.returncode != 0 or not line:
                    tried.append({"knobs": knobs, "why": "self-check failed (rc %d): %s" % (r.returncode, (r.stderr or r.stdout).strip()[-200:])})
                    break                                                # a failing opt-in is not tried again with other knobs
                d = json.loads(line[-1])
                tried.append({"knobs": knobs, "bit_identical": bool(d["ok"]), "base_ms": d["base_ms"], "feat_ms": d["feat_ms"]})
                if not d["ok"]:
                    break
                if d["feat_ms"] < 0.97 * d["base_ms"] and (best is None or d["feat_ms"] / d["base_ms"] < best[0]):
                    best = (d["feat_ms"] / d["base_ms"], knobs)
            except Exception as exc:                                     # noqa: BLE001 - timeout, missing file, bad JSON: the opt-in stays off
                tried.append({"knobs": knobs, "why": "self-check did not finish: %s" % str(exc)[:200]})
                break
        ok_all = all(t.get("bit_identical") for t in tried)
        out[feature] = {"used": bool(best) and ok_all, "tried": tried}
        if best and ok_all:
            os.environ[var] = "1"
            os.environ.update(best[1])
            out[feature]["knobs"] = best[1]
            out[feature]["ratio"] = best[0]
    # the path-tracer opt-ins were checked one by one; together they select kernel variants none of those runs launched: check the set
    pt_on = [(f, v) for f, v, _ in features[:3] if out.get(f, {}).get("used") and "tried" in out[f]]
    if len(pt_on) >= 2:
        combined = {"features": [f for f, _ in pt_on]}
        keep_all = False
        try:
            cmd = [sys.executable, os.path.join(ROOT, "tools", "selfcheck.py"), pt_on[0][0], "--config", args.config, "--mode", args.mode]
            for f, v in pt_on[1:]:
                cmd += ["--env", "%s=1" % v]
            for f, _ in pt_on:
                for k, val in out[f].get("knobs", {}).items():
                    cmd += ["--env", "%s=%s" % (k, val)]
            r = subprocess.run(cmd, capture_output=True, text=True, timeout=150)
            line = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
            if r.returncode == 0 and line:
                d = json.loads(line[-1])
                combined.update({"bit_identical": bool(d["ok"]), "base_ms": d["base_ms"], "feat_ms": d["feat_ms"]})
                keep_all = bool(d["ok"]) and d["feat_ms"] / d["base_ms"] < min(out[f]["ratio"] for f, _ in pt_on)
            else:
                combined["why"] = "self-check failed (rc %d): %s" % (r.returncode, (r.stderr or r.stdout).strip()[-200:])
        except Exception as exc:                                         # noqa: BLE001
            combined["why"] = "self-check did not finish: %s" % str(exc)[:200]
        combined["used"] = keep_all
        out["pt_combined"] = combined
        if not keep_all:                                                 # keep only the single best one
            best_f = min(pt_on, key=lambda fv: out[fv[0]]["ratio"])[0]
            for f, v in pt_on:
                if f != best_f:
                    os.environ.pop(v, None)
                    for k in out[f].get("knobs", {}):
                        os.environ.pop(k, None)
                    out[f]["used"] = False
                    out[f]["why"] = "not better together with %s" % best_f
    return out


def try_pipelined_strips(argv, rank, world, child_timeout=240):
    """N > 1: the two-stream frame loop in strip mode (PTD_STRIP_PIPELINE=1: gated live-count mail) measured 20-34 %
    more frames/s than the serial loop, but it has not been soaked since the gated mail was written.  So it is tried FIRST, as a complete
    benchmark run in a child process group (one child per rank, its own rendezvous port, every device-side wait traps after 20 s);
    only if every rank's child exits cleanly does rank 0 print the child's JSON line.  Otherwise - fault, trap, timeout, on any rank -
    the caller carries on with the serial loop in this process, exactly the configuration of the committed scaling lines.
    The ranks agree through small files in the temp directory (same node; no CUDA, no process group in the parents).
    Returns the JSON line (rank 0) / "" (other ranks) when the children succeeded everywhere, else None."""
    t_start = time.time()
    key = "ptd_bench_sup_%s_%s" % (os.environ.get("MASTER_PORT", "0"), os.environ.get("TORCHELASTIC_RUN_ID", "none"))
    mine = os.path.join(tempfile.gettempdir(), "%s_%d" % (key, rank))
    try:
        os.remove(mine)
    except OSError:
        pass
    env = dict(os.environ)
    env["PTD_STRIP_PIPELINE"] = "1"
    env["MASTER_PORT"] = str(int(os.environ.get("MASTER_PORT", "29500")) + 17)
    env.pop("TORCHELASTIC_USE_AGENT_STORE", None)                        # the child group's rank 0 hosts its own store on the new port
    cmd = os.environ.get("PTD_BENCH_CHILD_CMD")                            # test hook: a stand-in for the child benchmark
    cmd = cmd.split() if cmd else [sys.executable, os.path.abspath(__file__)] + list(argv)
    line, ok = "", False
    try:
        r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=child_timeout)
        js = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
        ok = r.returncode == 0 and (rank != 0 or bool(js))
        line = js[-1] if js else ""
        if not ok:
            sys.stderr.write("bench.py: rank %d: two-stream strip loop child failed (rc %d): %s\n" % (rank, r.returncode, (r.stderr or "").strip()[-300:]))
    except Exception as exc:                                              # noqa: BLE001 - timeout or spawn failure
        sys.stderr.write("bench.py: rank %d: two-stream strip loop child did not finish: %s\n" % (rank, str(exc)[:200]))
    with open(mine + ".tmp", "w") as f:
        f.write("1" if ok else "0")
    os.replace(mine + ".tmp", mine)
    deadline = t_start + child_timeout + 120
    votes = {}
    while len(votes) < world and time.time() < deadline:
        for r_ in range(world):
            if r_ in votes:
                continue
            pth = os.path.join(tempfile.gettempdir(), "%s_%d" % (key, r_))
            try:
                if os.path.getmtime(pth) >= t_start - 30:                  # not a leftover of an earlier run
                    votes[r_] = open(pth).read().strip() == "1"
            except OSError:
                pass
        if len(votes) < world:
            time.sleep(0.2)
    if len(votes) == world and all(votes.values()):
        return line if rank == 0 else ""
    return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    # f16: fp16 activation storage + kind::f16 MMAs, fp32 accumulate - measured error identical to tf32 (same 10-bit mantissa; DESIGN.md
    # section 3 table), half the bytes.  tf32 (fp32 storage) is the library / CLI default because of its range.
    ap.add_argument("--mode", default="f16", choices=["2xf16", "tf32", "f16", "3xtf32", "fp32"])
    ap.add_argument("--config", default="C3", choices=["C2", "C3", "C4", "C5"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e", default="auto", choices=["auto", "calls", "fused", "async"], help="host-pointer API of the e2e leg at N = 1: the reference's two call sites "
                    "(ptd_pt_render_host + ptd_dn_forward_host), the one-call frame (ptd_frame_host: the G-buffer is downloaded but never uploaded again), "
                    "or auto = the one-call frame if - and only if - it reproduces the two-call path bit for bit on this box, else the two calls")
    ap.add_argument("--no-autotune", action="store_true", help="do not try the opt-in code paths (tools/selfcheck.py); N = 1 only")
    ap.add_argument("--no-pipeline", action="store_true", help="serial frame loop (path trace, then denoise, on one stream) instead of the two-stream loop")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    W, H = {"C2": (1280, 720), "C3": (1280, 720), "C4": (1920, 1080), "C5": (2560, 1440)}[args.config]
    if args.impl == "reference":
        args.steps = min(args.steps, 3)
        args.warmup = min(args.warmup, 1)
        return run_reference(args, W, H, args.config)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    autotune = None
    if world == 1 and not args.no_autotune and args.mode != "fp32":
        autotune = run_autotune(args)
    if world > 1 and not args.no_autotune and "PTD_STRIP_PIPELINE" not in os.environ:
        line = try_pipelined_strips(sys.argv[1:], int(os.environ.get("RANK", "0")), world)
        if line is not None:
            if line:
                print(line)
            return

    import torch
    import torch.distributed as dist
    from ai_path_tracer_denoiser_b200 import capi, weights

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if capi.device_count() < 1:
        raise SystemExit("bench.py: no CUDA device - the product has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        # NCCL prints its version banner to STDOUT when NCCL_DEBUG >= VERSION; the contract is ONE JSON line there, so fd 1 points at
        # stderr while the communicator comes up
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)
    P = W * H
    scene_path, desc = make_scene(args.config, W, H)
    sc = capi.Scene(path=scene_path)
    nfaces = sc.counts()[2]
    wfile = os.path.join(tempfile.gettempdir(), "ptd_bench_weights_%d.ptdw" % rank)
    weights.save_weights(weights.synthetic_state_dict(1234), wfile)
    # N > 1: ONE frame sequence, every frame tiled in row strips over the N GPUs (path tracer and denoiser), halo rows and live
    # counts exchanged by the kernels themselves over NVLink peer memory (ai_path_tracer_denoiser_b200/tiling.py)
    from ai_path_tracer_denoiser_b200 import tiling
    if world > 1 and args.mode == "fp32":
        raise SystemExit("bench.py: row strips need --mode tf32 or f16")
    pipe = tiling.StripPipeline(sc, wfile, rank, world, local, dist if world > 1 else None,
                                {"tf32": capi.DN_TF32, "f16": capi.DN_F16, "3xtf32": capi.DN_3XTF32, "fp32": capi.DN_FP32, "2xf16": capi.DN_2XF16}[args.mode])
    pt, dn = pipe.pt, pipe.dn
    Hp, Wp = dn.padded_size()
    # the frame loop: path trace of frame k + 1 overlaps the denoiser of frame k on a second stream (tiling.FrameLoop); the profiling
    # and host-API legs below use the serial single-stream form
    loop = tiling.FrameLoop(pipe, pipelined=not args.no_pipeline)        # (FrameLoop itself stays serial in strip mode, see tiling.py)
    stream = loop.s_dn
    sptr = C.c_void_p(stream.cuda_stream)
    gbuf = loop.gbuf[0]
    rgb = loop.rgb
    cam0 = sc.camera[0]
    frame0 = 0
    cams = [capi.frame_camera(cam0, frame0 + k) for k in range(args.warmup + args.steps + 2)]
    L = capi.lib()

    def step(k, reset):
        capi.check(L.ptd_pt_render(pt.h, cams[k].ctypes.data, 1, C.c_void_p(gbuf.data_ptr()), sptr), "ptd_pt_render")
        capi.check(L.ptd_dn_forward(dn.h, C.c_void_p(gbuf.data_ptr()), C.c_void_p(rgb.data_ptr()), 1 if reset else 0, sptr), "ptd_dn_forward")

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- warm-up, then K timed steps (device timing: CUDA events on the launch stream, max over ranks) ----
    for k in range(args.warmup):
        loop.frame(cams[k], k == 0)
    sync_all()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(loop.s_pt)                                  # first work of the timed region is frame 0's path trace
    for k in range(args.steps):
        loop.frame(cams[args.warmup + k], False)
    e1.record(loop.s_dn)                                  # last work is the last frame's denoiser
    sync_all()
    ms_total = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    launches_per_step = pt.launches() + dn.launches()
    if world > 1:
        t = torch.tensor([ms_total], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_step = ms_total / args.steps
    fps = args.steps / (ms_total * 1e-3)                 # one frame sequence, whatever N is (strong scaling)

    # ---- per-launch device times of one more step (same stream, CUDA events between launches) ----
    pt.profile(True)
    dn.profile(True)
    prof = []
    for rep in range(3):
        step(args.warmup + args.steps, False)
        torch.cuda.synchronize()
        prof.append((pt.launch_times(), dn.launch_times()))
    pt.profile(False)
    dn.profile(False)
    pt_ms = np.mean([p[0] for p in prof], axis=0)
    dn_named = prof[-1][1]
    dn_ms = np.mean([[m for _, m in p[1]] for p in prof], axis=0)
    live, run = pt.live_counts()
    live_local, run_local = list(live), run
    if world > 1:                                        # frame-wide live counts = sum over the strips
        t = torch.tensor(live, device="cuda", dtype=torch.int64)
        dist.all_reduce(t)
        live = [int(v) for v in t.tolist()]
        run = max(b + 1 for b in range(len(live)) if live[b] > 0)
    Pl = pt.P                                            # pixels of this rank's strip
    peaks = measured_peaks()
    table = layer_table(pipe.dn_rows[1] if world > 1 else Hp, Wp)      # rank 0's strip when tiled
    conv_idx = [i for i, (n, _) in enumerate(dn_named) if n in table]
    conv_ms = float(sum(dn_ms[i] for i in conv_idx))
    conv_flops = sum(table[dn_named[i][0]][0] for i in conv_idx)
    conv_bytes = sum(table[dn_named[i][0]][1] for i in conv_idx) * (0.5 if args.mode == "f16" else 1.0)   # fp16 activations: 2 B per element
    other_dn_ms = float(sum(dn_ms)) - conv_ms
    pt_total_ms = float(sum(pt_ms))
    trace_ms, shade_ms = float(sum(pt_ms[0::2])), float(sum(pt_ms[1::2]))             # launch order: pt_trace, pt_shade per bounce
    # algorithmic bytes (reference AoS records): intersect 44 B read + 36 B written per live path (bounce 0 generates its
    # rays: 36 B only), shade + compaction 36 + 44 B read and 44 B written per survivor (<= per live path), G-buffer 28 P + 24 P
    ll = live_local[:run_local]
    trace_bytes = 36.0 * Pl + sum(80.0 * n for n in ll[1:]) + 16.0 * Pl
    shade_bytes = 36.0 * Pl + sum(80.0 * n for n in ll[1:]) + sum(44.0 * n for n in ll[1:]) + 12.0 * Pl + 24.0 * Pl
    pt_bytes = trace_bytes + shade_bytes
    # kind::tf32 issues at half the bf16 rate (no separate measured figure exists); kind::f16 at the bf16 rate
    tf32_peak = peaks["bf16"] / (1.0 if args.mode == "f16" else 2.0)
    conv_kernel = "conv_tc_kernel" if args.mode != "fp32" else "conv3x3_fp32"
    kernels = [
        dict(kernel=conv_kernel, launches=len(conv_idx), ms=conv_ms, share=conv_ms / (pt_total_ms + float(sum(dn_ms))),
             tflops=conv_flops / (conv_ms * 1e-3) / 1e12, gbs=conv_bytes / (conv_ms * 1e-3) / 1e9),
        dict(kernel="pt_trace", launches=len(pt_ms[0::2]), ms=trace_ms, share=trace_ms / (pt_total_ms + float(sum(dn_ms))),
             gbs=trace_bytes / (trace_ms * 1e-3) / 1e9, rays=int(sum(ll)), mrays_per_s=sum(ll) / (trace_ms * 1e-3) / 1e6),
        dict(kernel="pt_shade", launches=len(pt_ms[1::2]), ms=shade_ms, share=shade_ms / (pt_total_ms + float(sum(dn_ms))),
             gbs=shade_bytes / (shade_ms * 1e-3) / 1e9),
        dict(kernel="pack/pool/unpack", launches=len(dn_ms) - len(conv_idx), ms=other_dn_ms, share=other_dn_ms / (pt_total_ms + float(sum(dn_ms)))),
    ]
    if conv_ms >= pt_total_ms:
        ach = conv_flops / (conv_ms * 1e-3) / 1e12
        t_tensor, t_hbm = conv_flops / (tf32_peak * 1e12), conv_bytes / (peaks["hbm"] * 1e9)
        if args.mode == "tf32" and t_tensor >= t_hbm:
            roof = dict(bound="tensor", kernel=conv_kernel, achieved=ach, peak=tf32_peak, unit="TFLOP/s", frac=ach / tf32_peak, traffic=None)
        else:
            g = conv_bytes / (conv_ms * 1e-3) / 1e9
            roof = dict(bound="hbm", kernel=conv_kernel, achieved=g, peak=peaks["hbm"], unit="GB/s", frac=g / peaks["hbm"], traffic=None,
                        tensor_tflops=ach, tensor_frac_of_tf32_peak=ach / tf32_peak)
    else:
        g = trace_bytes / (trace_ms * 1e-3) / 1e9
        roof = dict(bound="hbm", kernel="pt_trace", achieved=g, peak=peaks["hbm"], unit="GB/s", frac=g / peaks["hbm"], traffic=None,
                    note="BVH traversal is latency/divergence bound, not HBM bound: see mrays_per_s in kernels[] and DESIGN.md")
    try:                                                 # DRAM traffic per launch of the dominant kernel, from the committed ncu capture
        tr = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        roof["traffic"] = tr[roof["kernel"]]["dram_bytes_per_launch"]
        roof["traffic_note"] = "%s; %s" % (tr[roof["kernel"]]["launch"], tr["source"])
    except Exception:
        pass
    roof["conv"] = dict(kernel=conv_kernel, ms=conv_ms, tflops=conv_flops / (conv_ms * 1e-3) / 1e12, tensor_peak_tflops=tf32_peak,
                        tensor_frac=conv_flops / (conv_ms * 1e-3) / 1e12 / tf32_peak, gbs=conv_bytes / (conv_ms * 1e-3) / 1e9,
                        hbm_frac=conv_bytes / (conv_ms * 1e-3) / 1e9 / peaks["hbm"])
    roof["peak_source"] = peaks["source"] + ("; tf32 peak = measured bf16 burst / 2" if roof["bound"] == "tensor" or "tensor_tflops" in roof else "")
    roof["per_layer_ms"] = {n: round(float(m), 4) for (n, _), m in zip(dn_named, dn_ms)}
    roof["per_bounce_ms"] = {"pt_trace": [round(float(m), 4) for m in pt_ms[0::2]], "pt_shade": [round(float(m), 4) for m in pt_ms[1::2]]}
    # What DOES bound pt_trace: the L1 data pipe (one wavefront per distinct 128-byte line per load).  Host-side model (ptd_bvh_probe_order:
    # the BVH4 traversal of pt_trace restated on the CPU over incoherent probe rays, 7 loads per node visit + 3 per triangle test) against the
    # time measured above for the secondary bounces; the ncu capture of bounce 1 (profiles/r01u_ncu_pt.txt) reads 83 % for this pipe.
    if nfaces and rank == 0 and len(pt_ms) >= 4:
        try:
            pr = (C.c_double * 8)()
            capi.check(L.ptd_bvh_probe_order(sc.h, 64000, 7, 4, pr), "ptd_bvh_probe_order")
            sec_rays = float(sum(ll[1:]))
            sec_ms = float(sum(pt_ms[2::2]))
            sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
            sms = torch.cuda.get_device_properties(local).multi_processor_count
            wf_per_clk_sm = sec_rays * pr[0] / (sec_ms * 1e-3 * sm_mhz * 1e6 * sms)
            roof["l1_model"] = dict(kernel="pt_trace (bounces >= 1)", wavefronts_per_ray_arrival_order=round(pr[0], 1), wavefronts_per_ray_binned=round(pr[1], 1),
                                    achieved_wavefronts_per_clk_per_sm=round(wf_per_clk_sm, 3), peak=1.0, frac=round(wf_per_clk_sm, 3),
                                    note="host-side model of the traversal on incoherent probe rays (the real bounce-1 rays are more coherent), not a hardware counter")
        except Exception as exc:                         # noqa: BLE001 - a diagnostic, never worth the benchmark
            roof["l1_model"] = {"error": str(exc)[:120]}
    roof["kernels"] = kernels

    # ---- end to end ----
    # N == 1: through the host-pointer C ABI, the calls the reference's runCuda() would make (G-buffer D2H, H2D again, frame D2H).
    # N > 1: every rank runs its strip of the frame and reads its rows of the denoised frame back to pinned host memory each step;
    #        the per-step input is the camera record (84 B, passed by value into the kernels).
    if world == 1:
        host_g = torch.empty(10 * P, dtype=torch.float32).pin_memory()
        host_rgb = torch.empty(3 * P, dtype=torch.float32).pin_memory()
        # Host-pointer APIs of the e2e leg, slowest to fastest: the reference's two call sites; ptd_frame_host (one blocking call, no re-upload
        # of the G-buffer); ptd_frame_submit / ptd_frame_wait (frame k + 1 submitted before frame k is awaited: its path trace overlaps the
        # denoiser and the PCIe copies of frame k).  The last two were written after round 1's GPU budget was spent, so `auto` uses the
        # fastest one that - here and now - returns exactly what the two call sites return over a 3-frame recurrent sequence.
        host_g2 = [host_g, torch.empty(10 * P, dtype=torch.float32).pin_memory()]
        host_rgb2 = [host_rgb, torch.empty(3 * P, dtype=torch.float32).pin_memory()]
        gp = [C.c_void_p(t.data_ptr()) for t in host_g2]
        rp = [C.c_void_p(t.data_ptr()) for t in host_rgb2]

        def two_calls(k, reset):
            capi.check(L.ptd_pt_render_host(pt.h, cams[k].ctypes.data, 1, gp[0]), "ptd_pt_render_host")
            capi.check(L.ptd_dn_forward_host(dn.h, gp[0], rp[0], 1 if reset else 0), "ptd_dn_forward_host")

        def fused(k, reset):
            capi.check(L.ptd_frame_host(pt.h, dn.h, cams[k].ctypes.data, 1, 1 if reset else 0, gp[0], rp[0]), "ptd_frame_host")

        def submit(k, reset, slot):
            capi.check(L.ptd_frame_submit(pt.h, dn.h, cams[k].ctypes.data, 1, 1 if reset else 0, gp[slot], rp[slot]), "ptd_frame_submit")

        def wait():
            capi.check(L.ptd_frame_wait(pt.h), "ptd_frame_wait")

        e2e_api, e2e_check = args.e2e, None
        if e2e_api == "auto":
            e2e_check = {}
            refs = []
            for k in range(3):
                two_calls(k, k == 0)
                refs.append((host_g2[0].clone(), host_rgb2[0].clone()))
            try:
                same = True
                for k in range(3):
                    host_g2[0].zero_(); host_rgb2[0].zero_()
                    fused(k, k == 0)
                    same = same and bool(torch.equal(refs[k][0], host_g2[0])) and bool(torch.equal(refs[k][1], host_rgb2[0]))
                e2e_check["ptd_frame_host"] = "bit-identical to the two call sites over 3 frames" if same else "differs from the two call sites: not used"
            except Exception as exc:                      # noqa: BLE001 - any failure of a new entry point falls back to the measured path
                same = False
                e2e_check["ptd_frame_host"] = "failed its self-check (%s): not used" % str(exc)[:200]
            ok_fused = same
            try:
                same = True
                for t in host_g2 + host_rgb2:
                    t.zero_()
                submit(0, True, 0)
                for k in (1, 2):
                    submit(k, False, k & 1)
                    wait()
                    same = same and bool(torch.equal(refs[k - 1][0], host_g2[(k - 1) & 1])) and bool(torch.equal(refs[k - 1][1], host_rgb2[(k - 1) & 1]))
                wait()
                same = same and bool(torch.equal(refs[2][0], host_g2[0])) and bool(torch.equal(refs[2][1], host_rgb2[0]))
                e2e_check["ptd_frame_submit/wait"] = "bit-identical to the two call sites over 3 frames" if same else "differs from the two call sites: not used"
            except Exception as exc:                      # noqa: BLE001
                same = False
                e2e_check["ptd_frame_submit/wait"] = "failed its self-check (%s): not used" % str(exc)[:200]
                for _ in range(2):                        # drain whatever is still in flight before the buffers are used again
                    try:
                        wait()
                    except Exception:                     # noqa: BLE001
                        break
                torch.cuda.synchronize()
            e2e_api = "async" if same else ("fused" if ok_fused else "calls")
        e2e_slot = [0]
        if e2e_api == "async":
            # one step = submit the next frame, then wait for the oldest one; `e2e_flush` completes the frame still in flight at the end
            def e2e_step(k, reset):
                first = e2e_slot[0] == 0
                submit(k, reset, e2e_slot[0] & 1)
                e2e_slot[0] += 1
                if not first:
                    wait()
            def e2e_flush():
                if e2e_slot[0] > 0:
                    wait()
                    e2e_slot[0] = 0
            h2d, d2h = 84, 52 * P
        elif e2e_api == "fused":
            def e2e_step(k, reset):
                fused(k, reset)
            e2e_flush = lambda: None                      # noqa: E731
            h2d, d2h = 84, 52 * P                        # the camera record in, G-buffer + frame out
        else:
            def e2e_step(k, reset):
                two_calls(k, reset)
            e2e_flush = lambda: None                      # noqa: E731
            h2d, d2h = 40 * P, 52 * P
    else:
        r0, nr = pipe.pt_rows
        host_rgb = [torch.empty(3, nr * W, dtype=torch.float32).pin_memory() for _ in range(2)]
        rgb3 = rgb.view(3, P)
        copied = [torch.cuda.Event(), torch.cuda.Event()]
        e2e_k = [0]
        def e2e_step(k, reset):
            # frame k is enqueued and its rows are read back behind it on the denoiser stream; the host then waits for frame k - 1's
            # rows, so every frame reaches pinned host memory and the device never idles while the host waits
            i = e2e_k[0] & 1
            loop.frame(cams[k], reset)
            with torch.cuda.stream(loop.s_dn):
                host_rgb[i].copy_(rgb3[:, r0 * W:(r0 + nr) * W], non_blocking=True)
                copied[i].record(loop.s_dn)
            if e2e_k[0] > 0:
                copied[i ^ 1].synchronize()
            e2e_k[0] += 1
        h2d, d2h = 84, 12 * P
    for k in range(3):
        e2e_step(k, k == 0)
    if world == 1:
        e2e_flush()
    sync_all()
    t0 = time.perf_counter()
    for k in range(args.steps):
        e2e_step(args.warmup + k, False)
    if world == 1:
        e2e_flush()                                      # async API: the last frame reaches host memory inside the timed region
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_fps = args.steps / e2e_s

    # ---- context for N > 1: the same GPUs as N independent frame sequences (no tiling, no coupling), aggregate frames/s ----
    replicas = None
    if world > 1:
        rpt = capi.PathTracer(sc, device=local)
        rdn = capi.Denoiser(wfile, H, W, device=local, flags={"tf32": capi.DN_TF32, "f16": capi.DN_F16, "3xtf32": capi.DN_3XTF32, "2xf16": capi.DN_2XF16}[args.mode])
        def rstep(k, reset):
            capi.check(L.ptd_pt_render(rpt.h, cams[k].ctypes.data, 1, C.c_void_p(gbuf.data_ptr()), sptr), "ptd_pt_render")
            capi.check(L.ptd_dn_forward(rdn.h, C.c_void_p(gbuf.data_ptr()), C.c_void_p(rgb.data_ptr()), 1 if reset else 0, sptr), "ptd_dn_forward")
        for k in range(3):
            rstep(k, k == 0)
        sync_all()
        r0e, r1e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        r0e.record(stream)
        nrep = min(args.steps, 50)
        for k in range(nrep):
            rstep(3 + k, False)
        r1e.record(stream)
        sync_all()
        t = torch.tensor([r0e.elapsed_time(r1e)], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        replicas = {"value": world * nrep / (float(t.item()) * 1e-3), "unit": "frames/s (aggregate of %d independent untiled frame sequences, one per GPU)" % world}

    out = {"metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
           "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
           "dtype": {"tf32": "tf32 conv operands, f32 accumulate/storage; f32 path trace", "f16": "f16 conv operands/activation storage, f32 accumulate, f32 frame; f32 path trace",
                     "3xtf32": "3xtf32 (hi/lo split) conv operands, f32 accumulate/storage; f32 path trace",
                     "2xf16": "f32-equivalent convs (contract mode, rel-L2 <= 1e-5 vs the fp32 reference): fp32 values as fp16 hi/lo pairs, 3 tcgen05 kind::f16 passes, f32 accumulate; f32 path trace",
                     "fp32": "f32"}[args.mode],
           "data": "synthetic",
           "config": {"workload": "%s: %s" % (args.config, desc), "frames": "camera pan phi_k = phi_0 + 0.002 k, recurrent hidden state carried",
                      "frame_loop": "two streams, double-buffered G-buffer: path trace of frame k+1 overlaps the denoiser of frame k" if loop.pipelined else "serial, one stream",
                      "triangles": nfaces, "live_paths_per_bounce": live[:run], "denoiser_padded": [Hp, Wp], "weights": "synthetic seed 1234 (no checkpoint ships)",
                      "l2": "per-frame working set (>1.5 GB of activations) exceeds the 126 MB L2; no explicit flush",
                      "parallelism": "1 GPU" if world == 1 else "each frame tiled in %d row strips (path tracer + denoiser), one strip per GPU; halo rows / live counts "
                                      "stored into the neighbours' memory by the kernels over NVLink (CUDA IPC peer pointers), no host or NCCL call per frame" % world,
                      "strip_rows_rank0": list(pipe.dn_rows)},
           "gpu_launches": launches_per_step * args.steps,
           "e2e": {"value": e2e_fps, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": e2e_s / args.steps * 1e3,
                   "api": {"async": "ptd_frame_submit + ptd_frame_wait (frame k + 1 submitted before frame k is awaited; camera in, G-buffer + denoised frame out to pinned host memory, every frame)",
                           "fused": "ptd_frame_host (one blocking call per frame: camera in, G-buffer + denoised frame out to host memory)",
                           "calls": "ptd_pt_render_host + ptd_dn_forward_host"}[e2e_api] if world == 1
                          else "ptd_pt_render + ptd_dn_forward per strip, frame rows read back to pinned host memory"},
           "roofline": roof, "clocks": clocks}
    if world == 1 and e2e_check:
        out["e2e"]["self_check"] = e2e_check
    if autotune is not None:
        out["config"]["autotune"] = autotune
    if world > 1 and getattr(pipe, "two_stream_ok", False):
        out["config"]["strip_loop"] = "PTD_STRIP_PIPELINE=1: two-stream loop on top of the gated live-count mail (DESIGN.md section 4); this whole run is the child " \
                                      "process group bench.py tries first - its line is only printed because every rank's child finished cleanly"
    if replicas:
        out["replicas"] = replicas
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        row_stride = max(1, H // 64) if nfaces else max(1, H // 256)
        pt_s, dn_s, rays = cpu_reference_frame(scene_path, cams[args.warmup], live[:run], H, W, threads, row_stride)
        out["cpu_baseline"] = {"value": 1.0 / (pt_s + dn_s), "unit": "frames/s", "cores": threads, "kind": "reference(path trace: oracle/_ref)+port(denoiser: oracle/dn_oracle.py)",
                               "sample": "brute-force first-bounce intersect of every %d-th row (%d rays x %d faces) scaled to this frame's %d live path-bounces -> %.1f s; one %dx%d torch-CPU forward -> %.2f s" % (
                                   row_stride, rays, nfaces, sum(live[:run]), pt_s, Hp, Wp, dn_s)}
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
