"""Runs a few frames of the bench workload (for ncu captures): python tools/run_frames.py [config] [frames] [mode]"""
import ctypes as C
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bench  # noqa: E402
from ai_path_tracer_denoiser_b200 import capi, weights  # noqa: E402

config = sys.argv[1] if len(sys.argv) > 1 else "C3"
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 2
mode = sys.argv[3] if len(sys.argv) > 3 else "tf32"
W, H = {"C2": (1280, 720), "C3": (1280, 720), "C4": (1920, 1080), "C5": (2560, 1440)}[config]
path, desc = bench.make_scene(config, W, H)
sc = capi.Scene(path=path)
pt = capi.PathTracer(sc)
wfile = os.path.join(tempfile.gettempdir(), "ptd_run_frames.ptdw")
weights.save_weights(weights.synthetic_state_dict(1234), wfile)
dn = capi.Denoiser(wfile, H, W, flags={"tf32": capi.DN_TF32, "f16": capi.DN_F16, "3xtf32": capi.DN_3XTF32, "fp32": capi.DN_FP32, "2xf16": capi.DN_2XF16}[mode])
g = torch.empty(10 * W * H, dtype=torch.float32, device="cuda")
rgb = torch.empty(3 * W * H, dtype=torch.float32, device="cuda")
L = capi.lib()
for k in range(frames):
    cam = capi.frame_camera(sc.camera[0], k)
    capi.check(L.ptd_pt_render(pt.h, cam.ctypes.data, 1, C.c_void_p(g.data_ptr()), None))
    capi.check(L.ptd_dn_forward(dn.h, C.c_void_p(g.data_ptr()), C.c_void_p(rgb.data_ptr()), 1 if k == 0 else 0, None))
torch.cuda.synchronize()
print("frames", frames, "live", pt.live_counts())
