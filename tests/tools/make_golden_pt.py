"""Generates scenes/*.txt (+ the small procedural OBJ) and tests/golden/pt_*.npz.

The golden vectors come from oracle/_ref (oracle/ref_driver.cu): the reference's own scene loader
(scene.cpp) and its own __host__ __device__ intersection / scatter functions compiled as host code, driven
by restated kernel loops.  Run in the build container only (needs /root/reference to build oracle/_ref)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from ai_path_tracer_denoiser_b200 import scenegen  # noqa: E402
from oracle import reflib, pt_oracle  # noqa: E402

W, H = 64, 48


def main():
    sc = os.path.join(ROOT, "scenes")
    scenegen.write_cornell(os.path.join(sc, "cornell_64x48.txt"), W, H, variant="diffuse")
    scenegen.write_cornell(os.path.join(sc, "cornell_specular_64x48.txt"), W, H, variant="specular")
    n = scenegen.write_obj(os.path.join(sc, "hall_small.obj"), "sponza", 1200)
    scenegen.write_mesh_scene(os.path.join(sc, "hall_64x48.txt"), "hall_small.obj", W, H, kind="sponza")
    scenegen.write_mesh_scene(os.path.join(sc, "hall_reflective_64x48.txt"), "hall_small.obj", W, H, kind="sponza", material="reflective")
    print("hall_small.obj triangles:", n)
    R = reflib.RefLib()
    for name, frame, sort in (("cornell_64x48", 0, False), ("cornell_specular_64x48", 0, False), ("cornell_specular_64x48", 7, True),
                              ("hall_64x48", 0, False), ("hall_reflective_64x48", 150, False)):
        s = R.load_scene(os.path.join(sc, name + ".txt"))
        A = R.scene_arrays(s)
        cam = pt_oracle.frame_camera(A["camera"][0], frame)
        R.set_camera(s, cam)
        r = R.cpu_render(s, sort_material=sort, trace=True)
        out = dict(geoms=A["geoms"], materials=A["materials"], faces=A["faces"], mesh_box=A["mesh_box"], camera_loaded=A["camera"],
                   camera=cam, depth=A["depth"], iterations=A["iterations"], frame=frame, sort=int(sort), tensor=r["tensor"],
                   image=r["image"], final_paths=r["final_paths"], counts=np.array([b["n"] for b in r["trace"]], np.int32))
        for b, t in enumerate(r["trace"]):
            out["paths%d" % b] = t["paths"]
            isx = t["isx"].copy()
            isx["pad"] = 0
            out["isx%d" % b] = isx
        tag = "%s_f%d%s" % (name, frame, "_sort" if sort else "")
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", "pt_%s.npz" % tag), **out)
        print(tag, out["counts"], "sum", r["sum_live"])


if __name__ == "__main__":
    main()
