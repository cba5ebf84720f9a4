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
