"""Debug aid: where the path trace and the denoiser of consecutive frames sit in time inside the ptd_frame_submit loop.
Needs a library built with -DPTD_FRAME_SPANS (PTD_LIBPTD=...); see DESIGN.md section 6."""
import ctypes, json, os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ai_path_tracer_denoiser_b200 import capi, scenegen, weights
mode = sys.argv[1] if len(sys.argv) > 1 else "2xf16"
lib = ctypes.CDLL(capi.LIB_PATH)
tmp = tempfile.mkdtemp()
path, _ = scenegen.make_config(tmp, "C3")
sc = capi.Scene(path=path)
pt = capi.PathTracer(sc, flags=0)
wfile = weights.save_weights(weights.synthetic_state_dict(1234), os.path.join(tmp, "w.ptdw"))
dn = capi.Denoiser(wfile, 720, 1280, flags={"f16": capi.DN_F16, "2xf16": capi.DN_2XF16, "tf32": capi.DN_TF32}[mode])
cams = [capi.frame_camera(sc.camera[0], k) for k in range(40)]
N = 24
for k in range(N):
    pt.frame_submit(dn, None, None, cam=cams[k], reset=(k == 0))
    if k >= capi.frame_slots() - 1:
        pt.frame_wait()
for _ in range(capi.frame_slots() - 1):
    pt.frame_wait()
out = (ctypes.c_float * 8)()
lib.ptd_debug_frame_spans(pt.h, out)
v = [round(x, 3) for x in out]
print(json.dumps({"mode": mode, "frame k-1": dict(zip(["pt0", "pt1", "dn0", "dn1"], v[:4])), "frame k": dict(zip(["pt0", "pt1", "dn0", "dn1"], v[4:]))}))
