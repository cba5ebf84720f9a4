"""CPU tests: the plain-C path-trace oracle (oracle/pt_oracle.c) against the golden vectors produced from the
reference's own host/device functions (tests/tools/make_golden_pt.py), the survey's known-answer vectors, and -
when oracle/_ref is present - the reference build itself on further seeded cases.  Bit-exact throughout."""
import ctypes as C
import glob
import os

import numpy as np
import pytest

from conftest import GOLDEN, SCENES
from oracle import pt_oracle, reflib

GOLD = sorted(glob.glob(os.path.join(GOLDEN, "pt_*.npz")))


def _same(a, b, skip=("pad",)):
    assert a.dtype.names == b.dtype.names and len(a) == len(b)
    for f in a.dtype.names:
        if f in skip:
            continue
        assert a[f].tobytes() == b[f].tobytes(), f


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[3:-4] for p in GOLD])
def test_oracle_matches_reference_golden(path):
    g = np.load(path)
    scene = dict(geoms=g["geoms"], materials=g["materials"], faces=g["faces"], mesh_box=g["mesh_box"], depth=int(g["depth"]))
    r = pt_oracle.render(scene, g["camera"], sort_material=bool(g["sort"]), trace=True)
    assert list(r["counts"][:len(g["counts"])]) == list(g["counts"])
    for b in range(len(g["counts"])):
        _same(r["trace"][b]["paths"], g["paths%d" % b])
        _same(r["trace"][b]["isx"], g["isx%d" % b])
    _same(r["final_paths"], g["final_paths"])
    assert r["tensor"].tobytes() == g["tensor"].tobytes()
    assert r["image"].tobytes() == g["image"].tobytes()


def test_frame_camera_matches_golden():
    g = np.load(os.path.join(GOLDEN, "pt_hall_reflective_64x48_f150.npz"))
    cam = pt_oracle.frame_camera(g["camera_loaded"][0], int(g["frame"]))
    assert cam.tobytes() == g["camera"].tobytes()


def test_known_answers():
    """SURVEY.md section 4 micro vectors (regenerated from the reference headers in the survey probes)."""
    L = pt_oracle.lib()
    assert L.pto_hash_seed(1, 12345, 8) == 939298829
    u = np.zeros(2, np.float32)
    L.pto_rng_draws(1, 12345, 8, 2, u.ctypes.data_as(C.c_void_p))
    assert np.allclose(u, [0.498972625, 0.907480478], rtol=0, atol=1e-9)
    d = np.zeros(3, np.float32)
    nrm = np.array([0, 1, 0], np.float32)
    L.pto_hemisphere(nrm.ctypes.data_as(C.c_void_p), 1, 12345, 8, d.ctypes.data_as(C.c_void_p))
    assert np.allclose(d, [0.388688952, 0.70637995, -0.591564238], rtol=0, atol=2e-7)
    f = np.zeros(1, reflib.FACE_DT)
    f["v"][0] = [[0, 0, -1], [1, 0, -1], [0, 1, -1]]
    f["n"][0] = [[0, 0, 1]] * 3
    o = np.array([.2, .2, 0], np.float32)
    dd = np.array([0, 0, -1], np.float32)
    ip, nn = np.zeros(3, np.float32), np.zeros(3, np.float32)
    t = L.pto_triangle(f.ctypes.data, o.ctypes.data, dd.ctypes.data, ip.ctypes.data, nn.ctypes.data)
    assert t == 1.0 and np.allclose(ip, [0.2, 0.6, -1.0])           # (sic) the reference's barycentric mix-up
    dd2 = np.array([0, 0, 1], np.float32)
    o2 = np.array([.2, .2, -2], np.float32)
    assert L.pto_triangle(f.ctypes.data, o2.ctypes.data, dd2.ctypes.data, ip.ctypes.data, nn.ctypes.data) == -1.0   # back face


def test_live_path_counts_cornell_400(tmp_path):
    """SURVEY.md section 8d KAT: Cornell 400x400 depth 8 live paths per bounce (sum 630 775 = 3.94 P)."""
    from ai_path_tracer_denoiser_b200 import scenegen
    g = np.load(os.path.join(GOLDEN, "pt_cornell_64x48_f0.npz"))
    cam = g["camera_loaded"].copy()
    cam["res"][0] = [400, 400]
    cam["pixlen"][0] = [2 * np.tan(np.float32(45 * (np.pi / 180))) / 400] * 2      # scene.cpp:143-150 for a square frame
    L = pt_oracle.lib()
    L.pto_camera_from_scene(cam.ctypes.data, C.c_float(45.0))
    scene = dict(geoms=g["geoms"], materials=g["materials"], faces=g["faces"], mesh_box=g["mesh_box"], depth=8)
    r = pt_oracle.render(scene, pt_oracle.frame_camera(cam[0], 0))
    assert list(r["counts"]) == [160000, 130707, 91222, 71609, 58252, 47695, 38998, 32292]
    assert r["sum_live"] == 630775


@pytest.mark.skipif(not reflib.available(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("scene,frame,res", [("cornell_specular_64x48", 123, (80, 56)), ("hall_reflective_64x48", 40, (72, 40))])
def test_oracle_matches_reference_build_live(scene, frame, res, capfd):
    R = reflib.RefLib()
    s = R.load_scene(os.path.join(SCENES, scene + ".txt"))
    A = R.scene_arrays(s)
    cam = A["camera"].copy()
    cam["res"][0] = res
    L = pt_oracle.lib()
    L.pto_camera_from_scene(cam.ctypes.data, C.c_float(float(cam["fov"][0][1])))
    cam = pt_oracle.frame_camera(cam[0], frame)
    R.set_camera(s, cam)
    ref = R.cpu_render(s, trace=True)
    ora = pt_oracle.render(A, cam, trace=True)
    assert [b["n"] for b in ref["trace"]] == [b["n"] for b in ora["trace"]]
    for a, b in zip(ref["trace"], ora["trace"]):
        _same(a["paths"], b["paths"])
        _same(a["isx"], b["isx"])
    assert ref["tensor"].tobytes() == ora["tensor"].tobytes()
