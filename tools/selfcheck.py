#!/usr/bin/env python
"""tools/selfcheck.py FEATURE [--config C3] [--mode f16] - validate and time ONE opt-in code path against the default one, in a process
of its own (bench.py's --autotune runs it as a subprocess, so a fault in an opt-in can never take the benchmark down).

    FEATURE     switch (read when the handle is created)      compared on the bench workload
    ray_sort    PTD_PT_RAY_SORT=1      coherent ray binning    G-buffer + live counts of 4 frames, bit for bit; ms per path-trace frame
    wide_lookback PTD_PT_WIDE_LOOKBACK=1 block-wide look-back  same
    smem_stack  PTD_PT_SMEM_STACK=1    traversal stack in shared memory: same
    pdl         PTD_DN_PDL=1           programmatic dependent launch of the convs: denoised frames of a 4-frame recurrence, bit for bit;
                                                               ms per denoiser forward

Prints ONE JSON line {"feature", "ok", "base_ms", "feat_ms", "detail"} and exits 0 when the comparison ran (ok tells the outcome),
non-zero on any error.  Both handles live in this process; the switch is set only while the second one is created."""
import argparse
import ctypes as C
import json
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

SWITCH = {"ray_sort": "PTD_PT_RAY_SORT", "wide_lookback": "PTD_PT_WIDE_LOOKBACK", "smem_stack": "PTD_PT_SMEM_STACK", "pdl": "PTD_DN_PDL"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("feature", choices=sorted(SWITCH))
    ap.add_argument("--config", default="C3")
    ap.add_argument("--mode", default="f16", choices=["tf32", "f16", "3xtf32", "2xf16"])
    ap.add_argument("--frames", type=int, default=12)
    ap.add_argument("--env", action="append", default=[], help="K=V set while the opt-in handle is created (tuning knobs of the feature)")
    args = ap.parse_args()
    knobs = dict(kv.split("=", 1) for kv in args.env)
    import torch
    from ai_path_tracer_denoiser_b200 import capi, scenegen, weights
    if capi.device_count() < 1:
        raise SystemExit("selfcheck: no CUDA device")
    torch.cuda.set_device(0)
    d = os.path.join(tempfile.gettempdir(), "ptd_bench_scenes_r%s" % os.environ.get("RANK", "0"))    # the directory bench.py uses
    os.makedirs(d, exist_ok=True)
    scene_path, _ = scenegen.make_config(d, args.config)
    sc = capi.Scene(path=scene_path)
    cam0 = sc.camera[0]
    W, H = int(cam0["res"][0]), int(cam0["res"][1])
    P = W * H
    cams = [capi.frame_camera(cam0, k) for k in range(args.frames + 4)]
    var = SWITCH[args.feature]
    os.environ.pop(var, None)
    stream = torch.cuda.Stream()
    sptr = C.c_void_p(stream.cuda_stream)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def timed(fn, n):
        for k in range(2):
            fn(k)
        torch.cuda.synchronize()
        e0.record(stream)
        for k in range(n):
            fn(2 + k)
        e1.record(stream)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    detail = {}
    if args.feature in ("ray_sort", "wide_lookback", "smem_stack"):
        base = capi.PathTracer(sc)
        os.environ[var] = "1"
        os.environ.update(knobs)
        feat = capi.PathTracer(sc)
        os.environ.pop(var, None)
        for k in knobs:
            os.environ.pop(k, None)
        ga = torch.zeros(10 * P, dtype=torch.float32, device="cuda")
        gb = torch.zeros(10 * P, dtype=torch.float32, device="cuda")
        ok = True
        for k in range(4):
            base.render(C.c_void_p(ga.data_ptr()), cam=cams[k], stream=sptr)
            feat.render(C.c_void_p(gb.data_ptr()), cam=cams[k], stream=sptr)
            torch.cuda.synchronize()
            same = bool(torch.equal(ga, gb)) and base.live_counts() == feat.live_counts()
            ok = ok and same
        detail["launches"] = [base.launches(), feat.launches()]
        for k in (4, 5):                                               # the host-pointer entry point bench.py's e2e leg uses (its own streams)
            ok = ok and base.render_host(cams[k]).tobytes() == feat.render_host(cams[k]).tobytes()
        base_ms = timed(lambda k: base.render(C.c_void_p(ga.data_ptr()), cam=cams[k], stream=sptr), args.frames)
        feat_ms = timed(lambda k: feat.render(C.c_void_p(gb.data_ptr()), cam=cams[k], stream=sptr), args.frames)
        base_ms2 = timed(lambda k: base.render(C.c_void_p(ga.data_ptr()), cam=cams[k], stream=sptr), args.frames)      # A-B-A: drift shows up here
        detail["base_ms_again"] = base_ms2
        base_ms = min(base_ms, base_ms2)
    else:
        flags = {"tf32": capi.DN_TF32, "f16": capi.DN_F16, "3xtf32": capi.DN_3XTF32, "2xf16": capi.DN_2XF16}[args.mode]
        wfile = os.path.join(tempfile.gettempdir(), "ptd_selfcheck_weights.ptdw")
        weights.save_weights(weights.synthetic_state_dict(1234), wfile)
        pt = capi.PathTracer(sc)
        gs = []
        for k in range(4):                                             # real G-buffers of the pan
            g = torch.zeros(10 * P, dtype=torch.float32, device="cuda")
            pt.render(C.c_void_p(g.data_ptr()), cam=cams[k], stream=sptr)
            gs.append(g)
        torch.cuda.synchronize()
        base = capi.Denoiser(wfile, H, W, flags=flags)
        os.environ[var] = "1"
        os.environ.update(knobs)
        feat = capi.Denoiser(wfile, H, W, flags=flags)
        os.environ.pop(var, None)
        for k in knobs:
            os.environ.pop(k, None)
        oa = torch.zeros(3 * P, dtype=torch.float32, device="cuda")
        ob = torch.zeros(3 * P, dtype=torch.float32, device="cuda")
        ok = True
        for k in range(4):
            base.forward(C.c_void_p(gs[k].data_ptr()), C.c_void_p(oa.data_ptr()), k == 0, stream=sptr)
            feat.forward(C.c_void_p(gs[k].data_ptr()), C.c_void_p(ob.data_ptr()), k == 0, stream=sptr)
            torch.cuda.synchronize()
            ok = ok and bool(torch.equal(oa, ob)) and bool(torch.isfinite(ob).all())
        g_host = gs[0].cpu().numpy().reshape(10, H, W)                 # the host-pointer entry point bench.py's e2e leg uses (legacy default stream)
        for k in range(2):
            ok = ok and base.forward_host(g_host, reset=(k == 0)).tobytes() == feat.forward_host(g_host, reset=(k == 0)).tobytes()
        base_ms = timed(lambda k: base.forward(C.c_void_p(gs[k % 4].data_ptr()), C.c_void_p(oa.data_ptr()), False, stream=sptr), 3 * args.frames)
        feat_ms = timed(lambda k: feat.forward(C.c_void_p(gs[k % 4].data_ptr()), C.c_void_p(ob.data_ptr()), False, stream=sptr), 3 * args.frames)
        base_ms2 = timed(lambda k: base.forward(C.c_void_p(gs[k % 4].data_ptr()), C.c_void_p(oa.data_ptr()), False, stream=sptr), 3 * args.frames)
        detail["base_ms_again"] = base_ms2
        base_ms = min(base_ms, base_ms2)
    print(json.dumps({"feature": args.feature, "switch": var, "ok": bool(ok), "base_ms": round(base_ms, 4), "feat_ms": round(feat_ms, 4), "knobs": knobs, "detail": detail}))


if __name__ == "__main__":
    main()
