"""Debug aid: average time a pt_shade block spends in each phase of a tile (globaltimer stamps of thread 0), C3 at 720p.
Needs the measurement build: make -C ai_path_tracer_denoiser_b200/csrc debug; PTD_LIBPTD=.../libptd_dbg.so python tools/shade_prof.py
PTD_PT_SHADE_TILED=1 profiles the one-tile-per-block kernel instead of the pipelined one."""
import ctypes, json, os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ai_path_tracer_denoiser_b200 import capi, scenegen
lib = ctypes.CDLL(capi.LIB_PATH)
tmp = tempfile.mkdtemp()
path, _ = scenegen.make_config(tmp, "C3")
sc = capi.Scene(path=path)
pt = capi.PathTracer(sc, flags=0)
cams = [capi.frame_camera(sc.camera[0], k) for k in range(20)]
for f in range(3):
    pt.render_host(cam=cams[f])
lib.ptd_debug_shade_prof(None, 1)
N = 10
for f in range(N):
    pt.render_host(cam=cams[3 + f])
buf = (ctypes.c_ulonglong * 16)()
lib.ptd_debug_shade_prof(buf, 0)
tiles = max(buf[15], 1)
tiled = os.environ.get("PTD_PT_SHADE_TILED", "0") not in ("", "0")
names = (["ticket + mail", "loads", "shade", "slowest warp", "look-back", "stores"] if tiled else
         ["ticket + prefetch issue + closing barrier", "wait for tile data", "registers + shade", "scan + publish + stage", "wait for previous tile's prefix", "stores"])
out = {"kernel": "pt_shade_tiled" if tiled else "pt_shade (pipelined)", "us_per_tile": {n: round(buf[i] / tiles / 1000.0, 3) for i, n in enumerate(names)},
       "tiles_per_frame": buf[15] / N, "sum_us_per_tile": round(sum(buf[i] for i in range(6)) / tiles / 1000.0, 3)}
print(json.dumps(out))
