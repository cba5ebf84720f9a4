"""CPU tests of the host side: the C-ABI library loads and exports every symbol include/ptd.h declares, the scene /
OBJ ingest reproduces the reference loader's records bit for bit, the orbit camera matches the oracle, and the
compute entry points fail loudly (no CPU fallback) when there is no CUDA device."""
import ctypes as C
import glob
import os
import re

import numpy as np
import pytest

from conftest import ROOT, SCENES
from ai_path_tracer_denoiser_b200 import capi, weights
from oracle import pt_oracle, reflib


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "ptd.h")).read()
    declared = set(re.findall(r"\b(ptd_[a-z0-9_]+)\s*\(", hdr)) - {"ptd_halo_fn", "ptd_status"}
    L = capi.lib()
    missing = [s for s in sorted(declared) if not hasattr(L, s)]
    assert not missing, missing
    assert declared == set(capi.EXPORTS)
    assert [L.ptd_sizeof(i) for i in range(7)] == [44, 36, 248, 76, 44, 84, 24]


@pytest.mark.skipif(not reflib.available(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("scene", sorted(os.path.basename(p) for p in glob.glob(os.path.join(SCENES, "*.txt"))))
def test_scene_loader_matches_reference_loader(scene):
    R = reflib.RefLib()
    path = os.path.join(SCENES, scene)
    ref = R.scene_arrays(R.load_scene(path))
    ours = capi.Scene(path=path).arrays()
    assert ours["depth"] == ref["depth"] and ours["iterations"] == ref["iterations"]
    for k in ("geoms", "materials", "faces", "mesh_box"):
        assert len(ours[k]) == len(ref[k]), k
        assert ours[k].tobytes() == ref[k].tobytes(), k
    a, b = ours["camera"][0], ref["camera"][0]
    for f in capi.CAM_DT.names:
        if f == "right":                       # NaN in both (scene.cpp:148 uses the still-zero view)
            assert np.isnan(a[f]).all() == np.isnan(b[f]).all()
        else:
            assert a[f].tobytes() == b[f].tobytes(), f


@pytest.mark.skipif(not reflib.available() or not os.path.isdir("/root/reference/Inference/scenes/Scenes"), reason="reference scenes not present")
def test_scene_loader_on_reference_shipped_scenes():
    """Every geometry-only scene file the reference ships (Inference/scenes/Scenes/*.txt without MESH)."""
    R = reflib.RefLib()
    n = 0
    for path in sorted(glob.glob("/root/reference/Inference/scenes/Scenes/*.txt")):
        if "MESH" in open(path, errors="replace").read():
            continue
        ref = R.scene_arrays(R.load_scene(path))
        ours = capi.Scene(path=path).arrays()
        for k in ("geoms", "materials"):
            assert ours[k].tobytes() == ref[k].tobytes(), (path, k)
        for f in ("res", "pos", "lookat", "view", "up", "fov", "pixlen"):
            assert ours["camera"][0][f].tobytes() == ref["camera"][0][f].tobytes(), (path, f)
        n += 1
    assert n >= 5


def test_orbit_camera_matches_oracle():
    sc = capi.Scene(path=os.path.join(SCENES, "cornell_64x48.txt"))
    cam = sc.camera[0]
    for frame in (0, 1, 150, 299):
        a = capi.frame_camera(cam, frame)
        b = pt_oracle.frame_camera(cam, frame)
        assert a.tobytes() == b.tobytes()


def test_errors_do_not_cross_the_boundary():
    L = capi.lib()
    h = C.c_void_p()
    assert L.ptd_scene_load(b"/nonexistent/scene.txt", C.byref(h)) == -2 and not h
    assert b"cannot open" in L.ptd_last_error()
    with pytest.raises(capi.PtdError):
        capi.Scene(path=os.path.join(ROOT, "README.md"))           # no CAMERA block


def test_no_cpu_fallback_without_a_device(tmp_path):
    if capi.device_count() > 0:
        pytest.skip("a CUDA device is present")
    sc = capi.Scene(path=os.path.join(SCENES, "cornell_64x48.txt"))
    with pytest.raises(capi.PtdError, match="no CPU fallback"):
        capi.PathTracer(sc)
    w = weights.save_weights(weights.synthetic_state_dict(1234), str(tmp_path / "w.ptdw"))
    with pytest.raises(capi.PtdError, match="no CPU fallback"):
        capi.Denoiser(w, 64, 64)


def test_strip_partition_covers_the_padded_frame():
    """Row strips of the multi-GPU denoiser: contiguous, multiples of 32, cover the padded frame, sizes differ by <= 32."""
    from ai_path_tracer_denoiser_b200 import capi
    for H in (720, 1080, 1440, 64, 400):
        Hp = (H + 31) // 32 * 32
        for n in (1, 2, 4, 8):
            if n > Hp // 32:
                with pytest.raises(capi.PtdError):
                    capi.strip_partition(H, n, 0)
                continue
            parts = [capi.strip_partition(H, n, i) for i in range(n)]
            assert parts[0][0] == 0 and parts[-1][0] + parts[-1][1] == Hp
            for (a0, an), (b0, bn) in zip(parts, parts[1:]):
                assert a0 + an == b0
            assert all(r % 32 == 0 and r > 0 for _, r in parts)
            assert max(r for _, r in parts) - min(r for _, r in parts) <= 32
    with pytest.raises(capi.PtdError):
        capi.strip_partition(720, 0, 0)


def test_compute_entry_points_fail_loudly_without_a_device(tmp_path):
    """No CPU fallback: on a box without a GPU every compute constructor must raise, not silently compute on the host."""
    from ai_path_tracer_denoiser_b200 import capi, weights
    if capi.device_count() > 0:
        pytest.skip("a CUDA device is present")
    sc = capi.Scene(path=os.path.join(SCENES, "cornell_64x48.txt"))
    with pytest.raises(capi.PtdError, match="no CPU fallback"):
        capi.PathTracer(sc)
    w = weights.save_weights(weights.synthetic_state_dict(1234), str(tmp_path / "w.ptdw"))
    with pytest.raises(capi.PtdError, match="no CPU fallback"):
        capi.Denoiser(w, 48, 64)
    with pytest.raises(capi.PtdError, match="no CPU fallback"):
        capi.Denoiser(w, 64, 64, strip=(0, 32))


def test_usemtl_extension_assigns_per_face_materials(tmp_path):
    """`USEMTL 1` in the MESH block (ignored by the reference's parser): mtllib / usemtl give every face its own material, appended
    behind the scene file's MATERIAL blocks; without the keyword the mesh keeps the single `material k` of the reference."""
    (tmp_path / "m.mtl").write_text("newmtl red\nKd 0.8 0.1 0.1\nKs 0.5 0.5 0.5\nNs 32\nillum 2\n\nnewmtl mirror\nKd 0.9 0.9 0.9\nKs 1 1 1\nillum 3\n\n"
                                    "newmtl glass\nKd 1 1 1\nKs 1 1 1\nNi 1.5\nillum 7\n\nnewmtl lamp\nKd 1 1 1\nKe 4 4 2\n")
    (tmp_path / "m.obj").write_text("mtllib m.mtl\nv 0 0 0\nv 1 0 0\nv 0 1 0\nv 1 1 0\nvn 0 0 1\nf 1//1 2//1 3//1\nusemtl red\nf 2//1 4//1 3//1\n"
                                    "usemtl mirror\nf 1//1 2//1 4//1\nusemtl glass\nf 1//1 4//1 3//1\nusemtl lamp\nf 2//1 3//1 4//1\nusemtl nosuch\nf 1//1 3//1 2//1\n")
    head = "MATERIAL 0\nRGB 1 1 1\nSPECEX 0\nSPECRGB 0 0 0\nREFL 0\nREFR 0\nREFRIOR 0\nEMITTANCE 5\n\n"
    tail = ("MATERIAL 1\nRGB .5 .5 .5\nSPECEX 0\nSPECRGB 0 0 0\nREFL 0\nREFR 0\nREFRIOR 0\nEMITTANCE 0\n\n"
            "CAMERA\nRES 32 32\nFOVY 45\nITERATIONS 1\nDEPTH 4\nFILE x\nEYE 0 0 5\nLOOKAT 0 0 0\nUP 0 1 0\n\n")
    mesh = "MESH 0\nPATH %s\nmaterial 1\nTRANS 0 0 0\nROTAT 0 0 0\nSCALE 1 1 1\n%s\n"
    (tmp_path / "plain.txt").write_text(head + mesh % (tmp_path / "m.obj", "") + tail)
    (tmp_path / "mtl.txt").write_text(head + mesh % (tmp_path / "m.obj", "USEMTL 1\n") + tail)     # MESH stands BEFORE material 1
    plain = capi.Scene(path=str(tmp_path / "plain.txt")).arrays()
    assert len(plain["materials"]) == 2 and list(plain["faces"]["mat"]) == [1] * 6
    a = capi.Scene(path=str(tmp_path / "mtl.txt")).arrays()
    assert len(a["materials"]) == 6                                   # 2 from the scene file + 4 from the MTL, in that order
    assert list(a["faces"]["mat"]) == [1, 2, 3, 4, 5, 1]              # before any usemtl / unknown name: the block's `material 1`
    assert a["faces"]["v"].tobytes() == plain["faces"]["v"].tobytes() and a["materials"][:2].tobytes() == plain["materials"].tobytes()
    red, mirror, glass, lamp = a["materials"][2:]
    assert np.allclose(red["color"], [0.8, 0.1, 0.1]) and red["specex"] == 32 and red["refl"] == 0 and red["refr"] == 0 and red["emit"] == 0
    assert mirror["refl"] == 1 and mirror["refr"] == 0 and np.allclose(mirror["speccolor"], 1)
    assert glass["refr"] == 1 and glass["ior"] == np.float32(1.5)
    assert lamp["emit"] == 4 and np.allclose(lamp["color"], [1, 1, 0.5])


def test_header_is_plain_c_and_links(tmp_path):
    """include/ptd.h must be usable from C (the boundary is a C ABI): compile a C99 translation unit against it and link libptd.so."""
    import subprocess
    src = tmp_path / "abi.c"
    src.write_text('#include "ptd.h"\n#include <stdio.h>\n'
                   'int main(void) { ptd_scene* s = 0; ptd_status rc = ptd_scene_load("/nonexistent/scene.txt", &s);\n'
                   '  printf("%d %d %d %s\\n", ptd_version(), ptd_sizeof(0), rc, ptd_last_error()); return rc == PTD_ERR_IO ? 0 : 1; }\n')
    exe = tmp_path / "abi"
    gcc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
    libdir = os.path.join(ROOT, "ai_path_tracer_denoiser_b200")
    subprocess.check_call([gcc, "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                           "-L", libdir, "-lptd", "-Wl,-rpath," + libdir])
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.split()[:3] == ["100", "44", "-2"] and "cannot open scene file" in out.stdout


def test_cli_usage_and_no_device_message():
    """`ptd_cli` keeps the reference's command line (main.cpp:50-56): usage text without arguments; without a GPU it fails loudly."""
    import subprocess
    exe = os.path.join(ROOT, "ai_path_tracer_denoiser_b200", "ptd_cli")
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 1 and "SCENEFILE.txt" in r.stdout
    r = subprocess.run([exe, "/nonexistent.txt"], capture_output=True, text=True)
    assert r.returncode == 2 and "cannot open scene file" in r.stderr
    if capi.device_count() == 0:
        r = subprocess.run([exe, os.path.join(SCENES, "cornell_64x48.txt")], capture_output=True, text=True)
        assert r.returncode == 2 and "no CPU fallback" in r.stderr


def _bvh_probe(scene, nrays, brute):
    out = (C.c_double * 8)()
    capi.check(capi.lib().ptd_bvh_probe(scene.h, nrays, 11, brute, out), "ptd_bvh_probe")
    return list(out)


@pytest.mark.parametrize("knobs", [{}, {"PTD_BVH_MAX_LEAF": "2"}, {"PTD_BVH_MAX_LEAF": "8", "PTD_BVH_SWEEP": "64"}, {"PTD_BVH_LEAF_COST": "0.6"}],
                         ids=["default", "maxleaf2", "maxleaf8-sweep", "sah-leaves"])
def test_bvh_traversal_equals_brute_force_on_the_host(knobs, monkeypatch, tmp_path):
    """The BVH pt_trace walks is new work (the reference loops over every face, pathtrace.cu:258-269): on the host, the same 4-wide
    layout and traversal rules must return exactly the brute-force (face, t) for every probe ray - on the shipped hall mesh and on a
    generated sponza-like mesh (thin columns, arches, large floor quads), for the default build and for the tuning knobs."""
    from ai_path_tracer_denoiser_b200 import scenegen
    for k in ("PTD_BVH_MAX_LEAF", "PTD_BVH_SWEEP", "PTD_BVH_LEAF_COST"):
        monkeypatch.delenv(k, raising=False)
    for k, v in knobs.items():
        monkeypatch.setenv(k, v)
    small = capi.Scene(path=os.path.join(SCENES, "hall_64x48.txt"))
    r = _bvh_probe(small, 4000, 4000)
    assert r[0] == 0 and r[7] > 0.5, r
    obj = str(tmp_path / "sponza_like.obj")
    scenegen.write_obj(obj, "sponza", 12000)
    big = capi.Scene(path=scenegen.write_mesh_scene(str(tmp_path / "s.txt"), obj, 64, 48, kind="sponza", material="diffuse"))
    r = _bvh_probe(big, 6000, 6000)
    assert r[0] == 0 and r[7] > 0.5, r
    assert r[3] + 2 <= 96                                              # traversal stack of pt_trace (PT_STACK)
    assert 1.0 < r[1] < 60 and r[2] < 60, r                              # a sane tree: tens of node visits / triangle tests per ray, not thousands


@pytest.mark.parametrize("knobs", [{}, {"PTD_BVH_MAX_LEAF": "1"}, {"PTD_BVH8_GREEDY": "1"}], ids=["dp-collapse", "dp-collapse-leaf1", "greedy-collapse"])
def test_bvh8q_traversal_equals_brute_force_on_the_host(knobs, monkeypatch, tmp_path):
    """BVH8q (csrc/ptd_bvh8.cpp: 8-wide, 8-bit quantised child boxes, 80-byte nodes, SAH-optimal collapse) - the data structure the next
    trace kernel is designed around - must be conservative: its host traversal (node groups / triangle masks, octant slot order)
    returns exactly the brute-force (face, t) for every probe ray, with fewer node visits than the 4-wide layout."""
    from ai_path_tracer_denoiser_b200 import scenegen
    for k in ("PTD_BVH_MAX_LEAF", "PTD_BVH_SWEEP", "PTD_BVH_LEAF_COST", "PTD_BVH8", "PTD_BVH8_GREEDY", "PTD_BVH8_CPRIM"):
        monkeypatch.delenv(k, raising=False)
    obj = str(tmp_path / "sponza_like.obj")
    scenegen.write_obj(obj, "sponza", 12000)
    big = capi.Scene(path=scenegen.write_mesh_scene(str(tmp_path / "s.txt"), obj, 64, 48, kind="sponza", material="diffuse"))
    small = capi.Scene(path=os.path.join(SCENES, "hall_64x48.txt"))
    ref = _bvh_probe(big, 6000, 0)
    for k, v in knobs.items():
        monkeypatch.setenv(k, v)
    monkeypatch.setenv("PTD_BVH8", "1")
    r = _bvh_probe(small, 4000, 4000)
    assert r[0] == 0 and r[7] > 0.5, r
    r = _bvh_probe(big, 6000, 6000)
    assert r[0] == 0 and r[7] == ref[7], (r, ref)                       # same rays, same hit fraction
    assert r[1] < ref[1] and r[3] <= 16, (r, ref)                       # fewer node visits per ray; a shallow stack of node groups
    assert r[6] > (3.0 if "PTD_BVH8_GREEDY" in knobs else 5.5)          # children per 8-wide node: the DP collapse fills them


def test_bvh_probe_argument_errors():
    out = (C.c_double * 8)()
    cornell = capi.Scene(path=os.path.join(SCENES, "cornell_64x48.txt"))     # no mesh
    assert capi.lib().ptd_bvh_probe(cornell.h, 10, 1, 0, out) == capi.ERR_STATE if hasattr(capi, "ERR_STATE") else capi.lib().ptd_bvh_probe(cornell.h, 10, 1, 0, out) < 0
    assert capi.lib().ptd_bvh_probe(None, 10, 1, 0, out) < 0


def test_ray_binning_estimate_on_the_host(tmp_path):
    """ptd_bvh_probe_order: grouping incoherent rays by (origin cell, direction octant) - what PTD_PT_RAY_SORT does on the device -
    must not cost more L1 wavefronts or warp steps per ray than arrival order, and on a mesh of some size it must save some."""
    from ai_path_tracer_denoiser_b200 import scenegen
    obj = str(tmp_path / "sponza_like.obj")
    scenegen.write_obj(obj, "sponza", 20000)
    sc = capi.Scene(path=scenegen.write_mesh_scene(str(tmp_path / "s.txt"), obj, 64, 48, kind="sponza", material="diffuse"))
    out = (C.c_double * 8)()
    capi.check(capi.lib().ptd_bvh_probe_order(sc.h, 64000, 3, 3, out), "ptd_bvh_probe_order")
    arrival_wf, binned_wf, arrival_steps, binned_steps, bins, rays = list(out)[:6]
    assert rays >= 63000 and 8 < bins <= 8 * 8 ** 3
    assert binned_wf < 0.9 * arrival_wf and binned_steps <= arrival_steps * 1.02, list(out)
    assert capi.lib().ptd_bvh_probe_order(sc.h, 64000, 3, 9, out) < 0          # cell bits out of range


def test_cli_surface_and_error_behaviour_without_a_gpu(tmp_path):
    """ptd_cli keeps the reference's command line (`exe SCENEFILE.txt`, main.cpp:50-56): usage without arguments, loud failures for a
    missing scene / unknown mode, and - on a box without a CUDA device - a refusal to render instead of a CPU fallback."""
    import subprocess
    cli = os.path.join(ROOT, "ai_path_tracer_denoiser_b200", "ptd_cli")
    if not os.path.exists(cli):
        capi.lib()                                                    # builds libptd.so + ptd_cli in-tree
    r = subprocess.run([cli], capture_output=True, text=True)
    assert r.returncode == 1 and "SCENEFILE.txt" in r.stdout + r.stderr
    r = subprocess.run([cli, str(tmp_path / "nope.txt")], capture_output=True, text=True)
    assert r.returncode == 2 and "cannot open scene file" in r.stdout + r.stderr
    r = subprocess.run([cli, os.path.join(SCENES, "cornell_64x48.txt"), "--mode", "bogus"], capture_output=True, text=True)
    assert r.returncode == 1 and "unknown --mode" in r.stdout + r.stderr
    if capi.device_count() == 0:
        r = subprocess.run([cli, os.path.join(SCENES, "cornell_64x48.txt"), "--frames", "1"], capture_output=True, text=True)
        assert r.returncode == 2 and "no CPU fallback" in r.stdout + r.stderr


def test_python_constants_match_the_header():
    """capi.py mirrors include/ptd.h by hand: every flag / mode / status value must agree with the header's enums."""
    hdr = open(os.path.join(ROOT, "include", "ptd.h")).read()
    vals = {m.group(1): int(m.group(2)) for m in re.finditer(r"^\s*(PTD_[A-Z0-9_]+) = (-?\d+)u?[, ]", hdr, re.M)}      # enumerator lines only
    pairs = {"PTD_PT_SORT_MATERIAL": capi.PT_SORT_MATERIAL, "PTD_PT_TRACE": capi.PT_TRACE, "PTD_PT_NO_BVH": capi.PT_NO_BVH,
             "PTD_PT_KEEP_TERMINATED": capi.PT_KEEP_TERMINATED, "PTD_PT_GATED_MAIL": capi.PT_GATED_MAIL, "PTD_PT_RAY_SORT": capi.PT_RAY_SORT,
             "PTD_DN_FP32": capi.DN_FP32, "PTD_DN_TF32": capi.DN_TF32, "PTD_DN_3XTF32": capi.DN_3XTF32, "PTD_DN_F16": capi.DN_F16,
             "PTD_DN_FP32_BATCH_STATS": capi.DN_FP32_BATCH_STATS}
    for name, py in pairs.items():
        assert vals[name] == py, name
    pt_flags = [v for k, v in vals.items() if k.startswith("PTD_PT_")]
    assert len(set(pt_flags)) == len(pt_flags) and all(v & (v - 1) == 0 for v in pt_flags)       # distinct single bits
