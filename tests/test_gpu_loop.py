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


@pytest.mark.parametrize("mode", ["experimental-frame-host"])
def test_frame_host_equals_the_two_host_calls(tmp_path, mode):
    """ptd_frame_host (runCuda()'s body as one call: the G-buffer stays on the device) == ptd_pt_render_host followed by
    ptd_dn_forward_host, bit for bit, with and without the optional host copy of the G-buffer, recurrent state carried."""
    from ai_path_tracer_denoiser_b200 import capi, weights
    if capi.device_count() < 1:
        pytest.fail("no CUDA device")
    wfile = weights.save_weights(weights.synthetic_state_dict(1234), str(tmp_path / "w.ptdw"))
    sc = capi.Scene(path=os.path.join(SCENES, "hall_64x48.txt"))
    sc.set_resolution(160, 96)
    cams = [capi.frame_camera(sc.camera[0], k) for k in range(4)]
    pt_a, dn_a = capi.PathTracer(sc), capi.Denoiser(wfile, 96, 160, flags=capi.DN_TF32)
    pt_b, dn_b = capi.PathTracer(sc), capi.Denoiser(wfile, 96, 160, flags=capi.DN_TF32)
    for k, cam in enumerate(cams):
        g_ref = pt_a.render_host(cam)
        rgb_ref = dn_a.forward_host(g_ref, reset=(k == 0))
        g, rgb = pt_b.frame_host(dn_b, cam=cam, reset=(k == 0), want_gbuffer=(k % 2 == 0))
        if g is not None:
            assert g.tobytes() == g_ref.tobytes(), k
        assert rgb.tobytes() == rgb_ref.tobytes(), k


@pytest.mark.parametrize("mode", ["experimental-frame-submit-wait"])
def test_frame_submit_wait_equals_the_two_host_calls(tmp_path, mode):
    """ptd_frame_submit / ptd_frame_wait (two frames in flight while the next is submitted) == ptd_pt_render_host + ptd_dn_forward_host,
    bit for bit over a recurrent sequence, with and without the host copy of the G-buffer; call-order errors are reported, not ignored."""
    from ai_path_tracer_denoiser_b200 import capi, weights
    if capi.device_count() < 1:
        pytest.fail("no CUDA device")
    wfile = weights.save_weights(weights.synthetic_state_dict(1234), str(tmp_path / "w.ptdw"))
    sc = capi.Scene(path=os.path.join(SCENES, "hall_64x48.txt"))
    sc.set_resolution(160, 96)
    n = 7
    cams = [capi.frame_camera(sc.camera[0], k) for k in range(n)]
    pt_a, dn_a = capi.PathTracer(sc), capi.Denoiser(wfile, 96, 160, flags=capi.DN_TF32)
    ref = []
    for k, cam in enumerate(cams):
        g = pt_a.render_host(cam)
        ref.append((g, dn_a.forward_host(g, reset=(k == 0))))
    pt_b, dn_b = capi.PathTracer(sc), capi.Denoiser(wfile, 96, 160, flags=capi.DN_TF32)
    with pytest.raises(capi.PtdError):
        pt_b.frame_wait()                                              # nothing in flight
    rgb = [np.zeros((3, 96, 160), np.float32) for _ in range(n)]
    gb = [np.zeros((10, 96, 160), np.float32) if k % 3 != 2 else None for k in range(n)]
    pt_b.frame_submit(dn_b, rgb[0], gb[0], cam=cams[0], reset=True)
    # while a frame is in flight it owns both handles' state: every other entry point reports that instead of racing with it
    with pytest.raises(capi.PtdError, match="in flight"):
        dn_b.forward_host(ref[0][0], reset=False)
    with pytest.raises(capi.PtdError, match="in flight"):
        pt_b.live_counts()
    with pytest.raises(capi.PtdError, match="in flight"):
        dn_b.dump_hidden(0)
    slots = capi.frame_slots()
    assert slots == 3
    done = 0

    def wait_and_check():
        nonlocal done
        pt_b.frame_wait()                                              # the oldest frame in flight is complete
        assert rgb[done].tobytes() == ref[done][1].tobytes(), done
        if gb[done] is not None:
            assert gb[done].tobytes() == ref[done][0].tobytes(), done
        done += 1

    for k in range(1, n):
        pt_b.frame_submit(dn_b, rgb[k], gb[k], cam=cams[k])
        if k == slots - 1:
            with pytest.raises(capi.PtdError, match="in flight"):
                pt_b.frame_submit(dn_b, rgb[k], gb[k], cam=cams[k])   # one frame more than there are slots
        if k >= slots - 1:
            wait_and_check()
    while done < n:
        wait_and_check()
    with pytest.raises(capi.PtdError):
        pt_b.frame_wait()
    # a frame may stay on the device (no host pointer at all), and the device timer brackets a run of frames
    pt_b.frame_timer_start()
    for k in range(4):
        pt_b.frame_submit(dn_b, None, None, cam=cams[k], reset=(k == 0))
        if k >= slots - 1:
            pt_b.frame_wait()
    ms = pt_b.frame_timer_stop()
    for _ in range(slots - 1):
        pt_b.frame_wait()
    assert 0.0 < ms < 5000.0
    # accumulation (iter > 1) needs the first-hit planes of iteration 1, which live in the other slot: reported, not silently wrong
    with pytest.raises(capi.PtdError, match="iter == 1"):
        pt_b.frame_submit(dn_b, rgb[0], None, cam=cams[0], iter=2)
    g3, rgb3 = pt_b.frame_host(dn_b, cam=cams[3], reset=True)           # the blocking call still works afterwards
    g_ref = pt_a.render_host(cams[3])
    assert g3.tobytes() == g_ref.tobytes() and rgb3.tobytes() == dn_a.forward_host(g_ref, reset=True).tobytes()
