"""Renders a few C3 frames through the C ABI (no denoiser): the workload of the ncu captures of pt_trace / pt_shade in profiles/."""
import os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ai_path_tracer_denoiser_b200 import capi, scenegen
path, _ = scenegen.make_config(tempfile.mkdtemp(), "C3")
sc = capi.Scene(path=path)
pt = capi.PathTracer(sc, flags=0)
for f in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    pt.render_host(cam=capi.frame_camera(sc.camera[0], f))
print("live counts", pt.live_counts())
