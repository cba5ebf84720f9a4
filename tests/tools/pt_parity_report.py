"""GPU-box diagnostic: field-by-field parity report of libptd.so against oracle A (the unmodified reference kernels,
oracle/_ref/libref_pt*.so) on the committed scenes.  Test infrastructure, not product."""
import glob
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from ai_path_tracer_denoiser_b200 import capi  # noqa: E402
from oracle import reflib  # noqa: E402


def ulps(a, b):
    a = np.ascontiguousarray(a, np.float32).view(np.int32).astype(np.int64)
    b = np.ascontiguousarray(b, np.float32).view(np.int32).astype(np.int64)
    a = np.where(a < 0, -(a & 0x7fffffff), a)
    b = np.where(b < 0, -(b & 0x7fffffff), b)
    return np.abs(a - b)


def report(name, a, b):
    out = []
    for f in a.dtype.names:
        if f == "pad":
            continue
        x, y = a[f], b[f]
        if x.tobytes() == y.tobytes():
            continue
        if x.dtype == np.float32:
            u = ulps(x, y).reshape(len(x), -1).max(axis=1)
            bad = np.nonzero(u)[0]
            out.append("%s: %d recs, max %d ulp, first %d" % (f, len(bad), u.max(), bad[0]))
        else:
            bad = np.nonzero((x != y).reshape(len(x), -1).any(axis=1))[0]
            out.append("%s: %d recs differ, first %d (%r vs %r)" % (f, len(bad), bad[0], x[bad[0]], y[bad[0]]))
    print("   %-22s %s" % (name, "; ".join(out) if out else "identical"))


def main():
    for path in sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "pt_*.npz"))):
        g = np.load(path)
        sort = bool(g["sort"])
        name = os.path.basename(path)[3:-4]
        R = reflib.RefLib("sort" if sort else "")
        devnull = os.open(os.devnull, os.O_WRONLY)
        saved = os.dup(1)
        os.dup2(devnull, 1)
        s = R.load_scene(os.path.join(ROOT, "scenes", name.split("_f")[0] + ".txt"))
        os.dup2(saved, 1)
        R.set_camera(s, g["camera"])
        ref = R.gpu_render(s, trace=True)
        a = dict(geoms=g["geoms"], materials=g["materials"], faces=g["faces"], mesh_box=g["mesh_box"], depth=int(g["depth"]), camera=g["camera"])
        sc = capi.Scene(arrays=a)
        pt = capi.PathTracer(sc, flags=(capi.PT_SORT_MATERIAL if sort else 0) | capi.PT_TRACE | capi.PT_KEEP_TERMINATED)
        tensor = pt.render_host()
        counts, run = pt.live_counts()
        print(name, "counts ours", counts[:run], "ref", [b["n"] for b in ref["trace"]])
        for b in range(min(run, len(ref["trace"]))):
            if counts[b] != ref["trace"][b]["n"]:
                print("   bounce %d: count mismatch, stopping" % b)
                break
            report("bounce %d paths" % b, pt.dump_paths(b), ref["trace"][b]["paths"])
            report("bounce %d isx" % b, pt.dump_intersections(b), ref["trace"][b]["isx"])
        report("final paths", pt.dump_final_paths(), ref["final_paths"])
        u = ulps(tensor, ref["tensor"])
        print("   tensor: %d of %d values differ, max %d ulp; image identical: %s" % (np.count_nonzero(u), u.size, u.max(), pt.dump_image().tobytes() == ref["image"].tobytes()))


if __name__ == "__main__":
    main()
