"""GPU parity tests of the denoiser hot path (HP-2), through the C ABI (libptd.so).

Checkers: tests/golden/dn_*.npz (outputs of the reference's own AutoEncoder, tests/tools/make_golden_dn.py) and
oracle/dn_oracle.py (pinned against the same golden files) on further seeded inputs.
Stated tolerances (outputs are O(1)):
  PTD_DN_FP32  (FFMA convs)                 max-abs <= 1e-4, rel-L2 <= 1e-5   - fp32 re-association only
  PTD_DN_TF32  (tcgen05 kind::tf32 convs)   max-abs <= 2e-2, rel-L2 <= 5e-3   - 10-bit-mantissa operands, fp32 accumulate
                                            (what libtorch itself does for convs on Ampere+ with cudnn.allow_tf32 = True)
  PTD_DN_3XTF32 (hi/lo split operands)      max-abs <= 3e-4, rel-L2 <= 1e-4   - hi*hi + hi*lo + lo*hi on the tensor cores, fp32 accumulate
                                            (measured 3.4e-5 / 1.0e-5: ~80x tighter than tf32, ~10x looser than FFMA - tests/tools/dn_accuracy.py)
  PTD_DN_F16   (fp16 storage, kind::f16)    same bound: fp16 has the same 10-bit mantissa, fp32 accumulate, fp32 output frame
"""
import os

import numpy as np
import pytest

from conftest import GOLDEN

pytestmark = pytest.mark.gpu

TOL = {"fp32": (1e-4, 1e-5), "2xf16": (1e-4, 1e-5), "3xtf32": (1e-4, 1e-5), "tf32": (2e-2, 5e-3), "f16": (2e-2, 5e-3)}


def _capi():
    from ai_path_tracer_denoiser_b200 import capi
    if capi.device_count() < 1:
        pytest.fail("no CUDA device visible to libptd.so - the product has no CPU fallback")
    return capi


@pytest.fixture(scope="module")
def wfile(tmp_path_factory):
    from ai_path_tracer_denoiser_b200 import weights
    return weights.save_weights(weights.synthetic_state_dict(1234), str(tmp_path_factory.mktemp("w") / "w.ptdw"))


def _err(y, ref):
    d = y.astype(np.float64) - ref.astype(np.float64)
    return np.abs(d).max(), np.sqrt((d * d).sum() / max((ref.astype(np.float64) ** 2).sum(), 1e-30))


def _mode(capi, mode):
    return {"fp32": capi.DN_FP32, "tf32": capi.DN_TF32, "f16": capi.DN_F16, "3xtf32": capi.DN_3XTF32, "2xf16": capi.DN_2XF16}[mode]


@pytest.mark.parametrize("mode", ["fp32", "tf32", "f16", "3xtf32", "2xf16"])
@pytest.mark.parametrize("name", ["dn_64x96", "dn_32x32"])
def test_matches_reference_model_golden(name, mode, wfile):
    capi = _capi()
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    _, H, W = g["x"][0].shape
    dn = capi.Denoiser(wfile, H, W, flags=_mode(capi, mode))
    for j in range(len(g["x"])):
        y = dn.forward_host(g["x"][j], reset=(j == 0))          # forward(x, j): hidden carried for j > 0
        ma, rl = _err(y, g["y"][j])
        assert ma <= TOL[mode][0] and rl <= TOL[mode][1], (j, ma, rl)
    hid = dn.dump_hidden(2)
    ma, rl = _err(hid, g["hidden3"][0] if g["hidden3"].ndim == 4 else g["hidden3"])
    assert ma <= 4 * TOL[mode][0], ma


@pytest.mark.parametrize("mode", ["fp32", "tf32", "f16", "3xtf32", "2xf16"])
def test_pad_crop_recurrence_and_reset_vs_oracle(mode, wfile):
    """Non-/32 frame (zero pad bottom/right, crop; decision D3), 5-frame recurrence, then a reset."""
    capi = _capi()
    from ai_path_tracer_denoiser_b200 import weights
    from oracle.dn_oracle import DenoiserOracle, synthetic_gbuffer
    H, W = 72, 100
    O = DenoiserOracle(weights.synthetic_state_dict(1234))
    dn = capi.Denoiser(wfile, H, W, flags=_mode(capi, mode))
    assert dn.padded_size() == (96, 128)
    first = None
    for j in range(5):
        x = synthetic_gbuffer(H, W, seed=11, frame=j)
        ref = O.forward(x, reset=(j == 0))
        y = dn.forward_host(x, reset=(j == 0))
        assert y.shape == (3, H, W)
        ma, rl = _err(y, ref)
        assert ma <= TOL[mode][0] and rl <= TOL[mode][1], (j, ma, rl)      # no growth through the recurrence
        if j == 0:
            first = y
    for lvl in range(6):
        ma, _ = _err(dn.dump_hidden(lvl), O.hidden[lvl][0].numpy())
        assert ma <= 4 * TOL[mode][0], (lvl, ma)
    again = dn.forward_host(synthetic_gbuffer(H, W, seed=11, frame=0), reset=True)
    assert again.tobytes() == first.tobytes()                                # reset really zeroes the six hidden states


@pytest.mark.parametrize("mode", ["fp32", "tf32", "f16", "3xtf32", "2xf16"])
def test_full_size_720p_properties(mode, wfile):
    """BASELINE size (720p -> 736x1280 padded): finite, deterministic, translation-consistent in the interior
    (a frame shifted by 32 px gives the shifted output away from the borders: the net is fully convolutional)."""
    capi = _capi()
    from oracle.dn_oracle import synthetic_gbuffer
    H, W = 720, 1280
    dn = capi.Denoiser(wfile, H, W, flags=_mode(capi, mode))
    x = synthetic_gbuffer(H, W, seed=3)
    y1 = dn.forward_host(x, reset=True)
    y2 = dn.forward_host(x, reset=True)
    assert np.isfinite(y1).all() and y1.tobytes() == y2.tobytes()
    xs = np.zeros_like(x)
    xs[:, :, 32:] = x[:, :, :-32]
    ys = dn.forward_host(xs, reset=True)
    a, b = ys[:, 300:420, 32 + 300:32 + 900], y1[:, 300:420, 300:900]
    assert np.abs(a - b).max() <= 2 * TOL[mode][0]


@pytest.fixture(params=["tiled", "replicated"])
def repl_levels(request, monkeypatch):
    """All levels tiled (PTD_DN_REPL_LEVEL=6) or levels >= 1/8 resolution replicated on every strip (3, the default; read at create)."""
    if request.param == "replicated":
        monkeypatch.setenv("PTD_DN_REPL_LEVEL", "3")
    else:
        monkeypatch.setenv("PTD_DN_REPL_LEVEL", "6")                 # every level tiled
    return request.param


@pytest.mark.parametrize("mode", ["tf32", "f16", "3xtf32", "2xf16"])
@pytest.mark.parametrize("nstrips", [2, 3])
def test_row_strips_equal_the_full_frame(nstrips, mode, wfile, repl_levels):
    """Multi-GPU tiling on ONE device: the frame cut into row strips (32-row aligned, uneven), each strip a handle with its
    own arena, halo rows stored into the neighbour's apron by the conv epilogues + device-side flags.  Result must be
    bit-identical to the single-handle forward over 4 frames (recurrent state carried, then reset)."""
    capi = _capi()
    import torch
    from oracle.dn_oracle import synthetic_gbuffer
    H, W = 150, 100                                   # padded 160 x 128: 5 groups of 32 rows
    full = capi.Denoiser(wfile, H, W, flags=_mode(capi, mode))
    parts = [capi.strip_partition(H, nstrips, i) for i in range(nstrips)]
    strips = [capi.Denoiser(wfile, H, W, flags=_mode(capi, mode), strip=p) for p in parts]
    infos = [s.export_info() for s in strips]
    for i, s in enumerate(strips):
        s.connect(infos, i)
    g = torch.empty(10 * H * W, dtype=torch.float32, device="cuda")
    out = torch.zeros(3 * H * W, dtype=torch.float32, device="cuda")
    for j, reset in enumerate([True, False, False, True, False]):
        x = synthetic_gbuffer(H, W, seed=5, frame=j)
        ref = full.forward_host(x, reset=reset)
        g.copy_(torch.from_numpy(x).reshape(-1))
        out.zero_()
        capi.Denoiser.forward_group(strips, [g.data_ptr()] * nstrips, [out.data_ptr()] * nstrips, reset)
        torch.cuda.synchronize()
        y = out.cpu().numpy().reshape(3, H, W)
        assert y.tobytes() == ref.tobytes(), (j, float(np.abs(y - ref).max()))


def test_multi_process_strips_bit_exact_when_two_gpus():
    """One process per GPU (torchrun), strips connected through CUDA IPC: tools/check_strips_multi.py.  Needs >= 2 GPUs."""
    capi = _capi()
    if capi.device_count() < 2:
        pytest.skip("needs two GPUs (runs under `gpurun --gpus 2`)")
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29533", os.path.join(root, "tools", "check_strips_multi.py")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count("BIT-EXACT") == 2


@pytest.mark.parametrize("gated", ["1", "0"], ids=["two-streams-gated-mail", "one-stream"])
def test_multi_process_frame_submit_on_strips_when_two_gpus(gated):
    """ptd_frame_submit / ptd_frame_wait on row-strip handles, one process per GPU (tools/check_frame_strips.py): every rank's rows of the
    G-buffer and of the denoised frame reach its host buffers bit-identical to the untiled frames, in the contract mode, with the path
    trace of frame k + 1 overlapping the denoiser of frame k (gated live-count mail) and without.  Needs >= 2 GPUs."""
    capi = _capi()
    if capi.device_count() < 2:
        pytest.skip("needs two GPUs (runs under `gpurun --gpus 2`)")
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29541", os.path.join(root, "tools", "check_frame_strips.py"), "320", "250", "7", "2xf16", gated],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count("BIT-EXACT") == 2


@pytest.mark.parametrize("mode", ["fp32", "tf32"])
@pytest.mark.parametrize("H,W", [(1, 1), (33, 31), (40, 200)])
def test_edge_sizes(H, W, mode, wfile):
    """Degenerate and ragged frame sizes: 1 x 1 (one 32 x 32 padded tile), one pixel past a 32-multiple, wide and flat."""
    capi = _capi()
    from ai_path_tracer_denoiser_b200 import weights
    from oracle.dn_oracle import DenoiserOracle, synthetic_gbuffer
    O = DenoiserOracle(weights.synthetic_state_dict(1234))
    dn = capi.Denoiser(wfile, H, W, flags=_mode(capi, mode))
    for j in range(2):
        x = synthetic_gbuffer(H, W, seed=4, frame=j)
        ma, rl = _err(dn.forward_host(x, reset=(j == 0)), O.forward(x, reset=(j == 0)))
        assert ma <= TOL[mode][0] and rl <= TOL[mode][1], (j, ma, rl)


def test_errors_are_reported_not_thrown(tmp_path, wfile):
    """Bad weight files / arguments come back as status codes with a message (the reference would throw or exit)."""
    capi = _capi()
    bad = tmp_path / "bad.ptdw"
    bad.write_bytes(b"NOPE" + b"\0" * 64)
    with pytest.raises(capi.PtdError, match="bad magic"):
        capi.Denoiser(str(bad), 32, 32)
    with pytest.raises(capi.PtdError, match="cannot open"):
        capi.Denoiser(str(tmp_path / "missing.ptdw"), 32, 32)
    trunc = tmp_path / "trunc.ptdw"
    trunc.write_bytes(open(wfile, "rb").read()[:4096])
    with pytest.raises(capi.PtdError, match="truncated|missing"):
        capi.Denoiser(str(trunc), 32, 32)
    with pytest.raises(capi.PtdError):
        capi.Denoiser(wfile, 64, 64, strip=(0, 48))          # strips must be multiples of 32 rows
    with pytest.raises(capi.PtdError, match="tensor-core"):
        capi.Denoiser(wfile, 64, 64, flags=capi.DN_FP32, strip=(0, 32))


def test_long_sequence_300_frames_is_stable(wfile):
    """BASELINE config 3: a 300-frame sequence with the recurrent state carried and never reset.  Output stays finite and bounded,
    and tracks the fp32 CPU oracle (checked every 60 frames) with no error growth."""
    capi = _capi()
    from ai_path_tracer_denoiser_b200 import weights
    from oracle.dn_oracle import DenoiserOracle, synthetic_gbuffer
    H, W = 64, 96
    O = DenoiserOracle(weights.synthetic_state_dict(1234))
    dn = capi.Denoiser(wfile, H, W, flags=capi.DN_TF32)
    errs = []
    for j in range(300):
        x = synthetic_gbuffer(H, W, seed=9, frame=j % 17)           # a short loop of inputs, a long recurrence
        ref = O.forward(x, reset=(j == 0))
        y = dn.forward_host(x, reset=(j == 0))
        if j % 60 == 59 or j == 0:
            assert np.isfinite(y).all() and np.abs(y).max() < 50
            errs.append(_err(y, ref))
    assert all(ma <= TOL["tf32"][0] and rl <= TOL["tf32"][1] for ma, rl in errs), errs
    assert errs[-1][1] <= 3 * errs[0][1] + 1e-4, errs                # no growth through the recurrence


@pytest.mark.parametrize("mode", ["experimental-traced-export"])
def test_batch_stats_mode_matches_traced_export_semantics(mode, wfile):
    """PTD_DN_FP32_BATCH_STATS: BatchNorm with the statistics of the current activations + hidden state zeroed every
    frame == the reference AutoEncoder in train() mode (golden made from the reference itself).  Stated tolerance: max-abs 2e-3 on
    O(1) outputs (the statistics of the 3 x 5-pixel bottleneck amplify fp32 reduction-order differences), rel-L2 2e-4."""
    capi = _capi()
    g = np.load(os.path.join(GOLDEN, "dn_traced_96x160.npz"))
    dn = capi.Denoiser(wfile, 96, 160, flags=capi.DN_FP32_BATCH_STATS)
    for j in range(len(g["x"])):
        y = dn.forward_host(g["x"][j], reset=True)
        err = np.abs(y - g["y"][j])
        assert err.max() <= 2e-3 and np.sqrt((err ** 2).sum() / (g["y"][j] ** 2).sum()) <= 2e-4, (j, err.max())
    with pytest.raises(capi.PtdError):
        capi.Denoiser(wfile, 96, 160, flags=capi.DN_FP32_BATCH_STATS, strip=(0, 32))      # cuda-core engines are not tiled


@pytest.mark.parametrize("mode", ["experimental-pdl-tf32", "experimental-pdl-f16"])
def test_programmatic_dependent_launch_changes_nothing(mode, wfile, monkeypatch):
    """PTD_DN_PDL=1 (opt-in): the convs are launched with programmatic stream serialization - the next layer's prologue overlaps the
    tail of the current one and griddepcontrol.wait orders the data.  Must be bit-identical to the plain launches over a recurrent
    sequence, at a size with both many-tile and few-tile layers."""
    capi = _capi()
    from oracle.dn_oracle import synthetic_gbuffer
    flags = capi.DN_TF32 if mode.endswith("tf32") else capi.DN_F16
    H, W = 200, 328
    xs = [synthetic_gbuffer(H, W, seed=21, frame=j) for j in range(4)]
    outs = {}
    for pdl in ("0", "1"):
        monkeypatch.setenv("PTD_DN_PDL", pdl)
        dn = capi.Denoiser(wfile, H, W, flags=flags)
        outs[pdl] = [dn.forward_host(x, reset=(j == 0)) for j, x in enumerate(xs)]
    for a, b in zip(outs["0"], outs["1"]):
        assert a.tobytes() == b.tobytes()
