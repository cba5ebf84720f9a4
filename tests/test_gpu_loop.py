"""GPU test of the frame loop (ai_path_tracer_denoiser_b200/tiling.py:FrameLoop): the two-stream pipelined loop (path trace of
frame k + 1 overlapping the denoiser of frame k, double-buffered G-buffer) must give bit-identical frames to the serial loop."""
import os

import numpy as np
import pytest

from conftest import SCENES

pytestmark = pytest.mark.gpu


def test_pipelined_loop_equals_serial_loop(tmp_path):
    import torch
    from ai_path_tracer_denoiser_b200 import capi, tiling, weights
    if capi.device_count() < 1:
        pytest.fail("no CUDA device")
    wfile = weights.save_weights(weights.synthetic_state_dict(1234), str(tmp_path / "w.ptdw"))
    sc = capi.Scene(path=os.path.join(SCENES, "hall_64x48.txt"))
    sc.set_resolution(160, 96)
    cams = [capi.frame_camera(sc.camera[0], k) for k in range(12)]
    frames = {}
    for pipelined in (False, True):
        pipe = tiling.StripPipeline(sc, wfile, 0, 1, 0)
        loop = tiling.FrameLoop(pipe, pipelined=pipelined)
        out = []
        for k, cam in enumerate(cams):
            loop.frame(cam, k == 0)
            if pipelined:                                  # keep the queue full: read frame k back while frame k + 1 is in flight
                with torch.cuda.stream(loop.s_dn):
                    out.append(loop.rgb.clone())
            else:
                loop.synchronize()
                out.append(loop.rgb.clone())
        loop.synchronize()
        torch.cuda.synchronize()
        frames[pipelined] = [o.cpu().numpy() for o in out]
    for k in range(len(cams)):
        assert frames[True][k].tobytes() == frames[False][k].tobytes(), k
    assert np.isfinite(frames[True][-1]).all() and np.abs(frames[True][-1]).max() > 0
