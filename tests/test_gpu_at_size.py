"""GPU parity tests AT THE BASELINE SIZES (BASELINE.json configs C2 / C3), through the C ABI (libptd.so).

  * HP-1, C2: the Cornell box at 1280 x 720, depth 8 - bit-exact against oracle A (the reference's own pathtrace.cu compiled
    verbatim, oracle/_ref/libref_pt.so, run on the same GPU): the PathSegment array entering every bounce (= after every
    thrust::partition, pathtrace.cu:505-506), every ShadeableIntersection, the final partition layout, dev_image and the
    10-plane G-buffer (pathtrace.cu:422-528).
  * HP-1, C3: the procedural Sponza-like mesh (261 k triangles) at 1280 x 720, depth 8, one frame of the pan - the same
    comparison; the reference brute-forces every face for every ray (pathtrace.cu:258-269), ~1.8e12 triangle tests, once.
  * HP-2: 720p (padded 736 x 1280), 3 recurrent frames, every conv engine against oracle/dn_oracle.py
    (recurrent_autoencoder_model.py:120-142 in torch fp32 on the CPU) at the tolerance stated for the mode.
"""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _capi():
    from ai_path_tracer_denoiser_b200 import capi
    if capi.device_count() < 1:
        pytest.skip("no CUDA device visible to libptd.so - the product has no CPU fallback")
    return capi


def _same(a, b, what, skip=("pad",)):
    assert len(a) == len(b), "%s: length %d vs %d" % (what, len(a), len(b))
    for f in a.dtype.names:
        if f in skip:
            continue
        if a[f].tobytes() != b[f].tobytes():
            x = a[f].view(np.uint32) if a[f].dtype != np.uint8 else a[f]
            y = b[f].view(np.uint32) if b[f].dtype != np.uint8 else b[f]
            bad = np.nonzero(np.any(x.reshape(len(a), -1) != y.reshape(len(b), -1), axis=1))[0]
            raise AssertionError("%s field %s: %d of %d records differ, first at %d: %r vs %r" % (what, f, len(bad), len(a), bad[0], a[f][bad[0]], b[f][bad[0]]))


def _bit_exact_vs_oracle_a(capi, scene_path, frame, flags=0):
    from oracle import reflib
    if not reflib.available(""):
        pytest.fail("oracle/_ref/libref_pt.so missing on the GPU box (build it with `make -C oracle ref` before gpurun)")
    R = reflib.RefLib("")
    s = R.load_scene(scene_path)
    sc = capi.Scene(path=scene_path)
    cam = capi.frame_camera(sc.camera[0], frame)
    sc.set_camera(cam)
    R.set_camera(s, cam)
    pt = capi.PathTracer(sc, flags=flags | capi.PT_TRACE | capi.PT_KEEP_TERMINATED)
    tensor = pt.render_host()
    counts, run = pt.live_counts()
    ref = R.gpu_render(s, trace=True)
    assert [b["n"] for b in ref["trace"]] == counts[:run]
    for b in range(run):
        _same(pt.dump_paths(b), ref["trace"][b]["paths"], "bounce %d paths" % b)
        _same(pt.dump_intersections(b), ref["trace"][b]["isx"], "bounce %d intersections" % b)
    _same(pt.dump_final_paths(), ref["final_paths"], "final partition layout")
    assert pt.dump_image().tobytes() == ref["image"].tobytes()
    assert tensor.tobytes() == ref["tensor"].tobytes()
    # Without the reject array (the product configuration) pt_shade stores a tile's survivors one pipeline step later: same frame, bit for bit.
    del pt
    pt2 = capi.PathTracer(sc, flags=flags | capi.PT_TRACE)
    assert pt2.render_host().tobytes() == ref["tensor"].tobytes()
    assert pt2.live_counts()[0][:run] == counts[:run]
    for b in range(run):
        _same(pt2.dump_paths(b), ref["trace"][b]["paths"], "deferred stores, bounce %d paths" % b)
        _same(pt2.dump_intersections(b), ref["trace"][b]["isx"], "deferred stores, bounce %d intersections" % b)
    assert pt2.dump_image().tobytes() == ref["image"].tobytes()
    del pt2
    pt3 = capi.PathTracer(sc, flags=flags)
    assert pt3.render_host().tobytes() == ref["tensor"].tobytes()
    assert pt3.live_counts()[0][:run] == counts[:run]
    return counts[:run], ref["ms"]


def test_c2_cornell_720p_bit_exact_vs_reference_kernels(tmp_path):
    """BASELINE config C2.  Live counts are also the survey's CPU-probe KAT up to borderline rays."""
    capi = _capi()
    from ai_path_tracer_denoiser_b200 import scenegen
    path, _ = scenegen.make_config(str(tmp_path), "C2")
    counts, _ = _bit_exact_vs_oracle_a(capi, path, 0)
    assert counts[0] == 921600 and len(counts) == 8
    for a, b in zip(counts, [921600, 423505, 294406, 231943, 189033, 155133, 126972, 104657]):
        assert abs(a - b) <= 0.002 * b


@pytest.mark.parametrize("opts", [{}, {"PTD_PT_RAY_SORT": "1", "PTD_PT_SMEM_STACK": "1"}, {"PTD_PT_RAY_SORT": "1", "PTD_PT_SMEM_STACK": "1", "PTD_PT_WIDE_LOOKBACK": "1"},
                                  ], ids=["default", "kernel-options", "kernel-options-tiled-shade"])
def test_c3_sponza_like_720p_bit_exact_vs_reference_kernels(tmp_path, opts, monkeypatch):
    """BASELINE config C3, frame 7 of the pan: 921 600 camera rays, 6.9 M path-bounces, ~1 800 shade tiles per bounce in the
    decoupled look-back, the BVH deciding which of the 261 k faces get the exact test.  Run with the default kernels and with every
    scheduling option bench.py's autotuner may switch on."""
    capi = _capi()
    from ai_path_tracer_denoiser_b200 import scenegen
    for k in ("PTD_PT_RAY_SORT", "PTD_PT_SMEM_STACK", "PTD_PT_WIDE_LOOKBACK", "PTD_PT_SHADE_TILED"):
        monkeypatch.delenv(k, raising=False)
    for k, v in opts.items():
        monkeypatch.setenv(k, v)
    path, _ = scenegen.make_config(str(tmp_path), "C3")
    counts, ref_ms = _bit_exact_vs_oracle_a(capi, path, 7)
    assert counts[0] == 921600 and len(counts) == 8 and counts[-1] > 400000
    print("reference kernels (brute force): %.0f ms for the frame" % ref_ms)


@pytest.fixture(scope="module")
def dn_720p_reference():
    """3 recurrent frames of the torch-fp32-CPU oracle at 720p, computed once for all modes."""
    from ai_path_tracer_denoiser_b200 import weights
    from oracle.dn_oracle import DenoiserOracle, synthetic_gbuffer
    H, W = 720, 1280
    O = DenoiserOracle(weights.synthetic_state_dict(1234))
    xs = [synthetic_gbuffer(H, W, seed=31, frame=j) for j in range(3)]
    refs = [O.forward(x, reset=(j == 0)) for j, x in enumerate(xs)]
    hidden = [h[0].numpy().copy() for h in O.hidden]
    return xs, refs, hidden


@pytest.mark.parametrize("mode", ["fp32", "2xf16", "3xtf32", "tf32", "f16"])
def test_denoiser_720p_three_recurrent_frames_vs_oracle(mode, dn_720p_reference, tmp_path):
    capi = _capi()
    from ai_path_tracer_denoiser_b200 import weights
    from test_gpu_dn import TOL, _err, _mode
    xs, refs, hidden = dn_720p_reference
    wfile = weights.save_weights(weights.synthetic_state_dict(1234), str(tmp_path / "w.ptdw"))
    dn = capi.Denoiser(wfile, 720, 1280, flags=_mode(capi, mode))
    assert dn.padded_size() == (736, 1280)
    errs = []
    for j, x in enumerate(xs):
        y = dn.forward_host(x, reset=(j == 0))
        errs.append(_err(y, refs[j]))
    print(mode, "max-abs / rel-L2 per frame:", ["%.2e / %.2e" % e for e in errs])
    for j, (ma, rl) in enumerate(errs):
        assert ma <= TOL[mode][0] and rl <= TOL[mode][1], (j, ma, rl)
    for lvl in range(6):
        ma, _ = _err(dn.dump_hidden(lvl), hidden[lvl])
        assert ma <= 4 * TOL[mode][0], (lvl, ma)
