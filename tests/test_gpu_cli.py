"""GPU test of the command-line surface: `ptd_cli SCENEFILE.txt` (the reference's command line, main.cpp:50-56) runs the headless
frame loop; its PFM dump of the denoised frame must equal what the C ABI gives through the ctypes mirror for the same frames."""
import os
import struct
import subprocess

import numpy as np
import pytest

from conftest import ROOT, SCENES

pytestmark = pytest.mark.gpu


def _read_pfm(path):
    with open(path, "rb") as f:
        assert f.readline().strip() == b"PF"
        W, H = map(int, f.readline().split())
        scale = float(f.readline())
        data = np.frombuffer(f.read(), "<f4" if scale < 0 else ">f4").reshape(H, W, 3)
    return data[::-1].transpose(2, 0, 1)          # bottom-up rows -> planar [3][H][W]


def _png_size(path):
    b = open(path, "rb").read()
    assert b[:8] == b"\x89PNG\r\n\x1a\n" and b[12:16] == b"IHDR"
    return struct.unpack(">II", b[16:24])


@pytest.mark.parametrize("roundtrip", [[], ["--host-roundtrip"], ["--serial"]], ids=["pipelined", "host-roundtrip", "serial"])
def test_cli_frame_loop_matches_the_c_abi(tmp_path, roundtrip):
    from ai_path_tracer_denoiser_b200 import capi, weights
    if capi.device_count() < 1:
        pytest.fail("no CUDA device")
    wfile = weights.save_weights(weights.synthetic_state_dict(1234), str(tmp_path / "w.ptdw"))
    scene = os.path.join(SCENES, "cornell_specular_64x48.txt")
    exe = os.path.join(ROOT, "ai_path_tracer_denoiser_b200", "ptd_cli")
    cmd = [exe, scene, "--weights", wfile, "--frames", "3", "--mode", "tf32", "--out", str(tmp_path / "f")] + roundtrip
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "3 frame(s)" in r.stdout
    sc = capi.Scene(path=scene)
    pt = capi.PathTracer(sc)
    dn = capi.Denoiser(wfile, 48, 64, flags=capi.DN_TF32)
    for k in range(3):
        cam = capi.frame_camera(sc.camera[0], k)
        ref = dn.forward_host(pt.render_host(cam), reset=(k == 0))
        got = _read_pfm(str(tmp_path / ("f_%04d_denoised.pfm" % k)))
        assert got.tobytes() == ref.tobytes(), k
        assert _png_size(str(tmp_path / ("f_%04d_denoised.png" % k))) == (64, 48)
        assert _png_size(str(tmp_path / ("f_%04d_1spp.png" % k))) == (64, 48)


def test_cli_path_trace_only_and_flags(tmp_path):
    exe = os.path.join(ROOT, "ai_path_tracer_denoiser_b200", "ptd_cli")
    r = subprocess.run([exe, os.path.join(SCENES, "hall_64x48.txt"), "--frames", "2", "--res", "96", "64", "--depth", "4", "--sort-material"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "96x64, depth 4" in r.stdout and "path trace only" in r.stdout
    r = subprocess.run([exe, os.path.join(SCENES, "hall_64x48.txt"), "--bogus"], capture_output=True, text=True)
    assert r.returncode == 1
