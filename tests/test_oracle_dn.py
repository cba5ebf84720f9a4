"""CPU tests: the denoiser oracle (oracle/dn_oracle.py) against golden outputs of the reference's own
AutoEncoder (tests/tools/make_golden_dn.py), plus the weight container round trip."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from ai_path_tracer_denoiser_b200 import weights
from oracle.dn_oracle import DenoiserOracle, synthetic_gbuffer

# Same torch CPU kernels on the same image give identical bits; a different host may reorder the
# convolution sums, so the stated tolerance is a few fp32 ulps of O(1) outputs.
ATOL = 2e-5


@pytest.mark.parametrize("name", ["dn_64x96", "dn_32x32"])
def test_oracle_matches_reference_model_golden(name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    O = DenoiserOracle(weights.synthetic_state_dict(1234))
    for j in range(len(g["x"])):
        y = O.forward(g["x"][j], reset=(j == 0))
        assert np.abs(y - g["y"][j]).max() <= ATOL
    assert np.abs(O.hidden[2][0].numpy() - g["hidden3"]).max() <= ATOL * 4


def test_oracle_batch_stats_mode_matches_reference_in_training_mode():
    """the module traced by convert_to_torchscript.py:26-30 is never put in eval mode - BatchNorm normalises with the
    statistics of its current input and j == 0 zeroes the hidden state on every call.  Golden = the reference AutoEncoder in
    train() mode (tests/tools/make_golden_dn.py); it must differ from the eval-mode forward, or the mode would be untested."""
    g = np.load(os.path.join(GOLDEN, "dn_traced_96x160.npz"))
    sd = weights.synthetic_state_dict(1234)
    O = DenoiserOracle(sd, batch_stats=True)
    for j in range(len(g["x"])):
        y = O.forward(g["x"][j], reset=True)
        assert np.abs(y - g["y"][j]).max() <= 5 * ATOL                 # batch statistics: one more reduction whose order may differ
    y_eval = DenoiserOracle(sd).forward(g["x"][0], reset=True)
    assert np.abs(y_eval - g["y"][0]).max() > 1e-2


def test_state_dict_layout_and_roundtrip(tmp_path):
    sd = weights.synthetic_state_dict(1234)
    assert len(sd) == 196
    assert sum(v.size for k, v in sd.items() if k.endswith((".weight", ".bias"))) == 1508181
    assert len(weights.conv_layers()) == 28
    p = weights.save_weights(sd, str(tmp_path / "w.ptdw"))
    back = weights.load_weights(p)
    assert len(back) == 168                                                        # 196 minus 28 num_batches_tracked
    for k, v in back.items():
        assert v.tobytes() == sd[k].tobytes()


def test_pad_crop_and_reset():
    O = DenoiserOracle(weights.synthetic_state_dict(1234))
    x = synthetic_gbuffer(40, 50, seed=3)
    y0 = O.forward(x, reset=True)
    y1 = O.forward(x, reset=False)
    y2 = O.forward(x, reset=True)
    assert y0.shape == (3, 40, 50)
    assert np.array_equal(y0, y2) and not np.array_equal(y0, y1)                   # hidden state matters
    assert O.hidden[0].shape == (1, 32, 64, 64) and O.hidden[5].shape == (1, 101, 2, 2)


def test_checkpoint_export_round_trip(tmp_path):
    """tools/export_weights.py: a `{'net': state_dict}` checkpoint as train.py writes it (train.py:108-112) -> PTDW -> same tensors."""
    import subprocess
    import sys
    import torch
    from ai_path_tracer_denoiser_b200 import weights
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sd = {k: torch.as_tensor(np.asarray(v)) for k, v in weights.synthetic_state_dict(7).items()}
    torch.save({"net": sd}, str(tmp_path / "model_3.pt"))
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "export_weights.py"), str(tmp_path / "model_3.pt"), str(tmp_path / "w.ptdw")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    back = weights.load_weights(str(tmp_path / "w.ptdw"))
    assert len(back) >= 28 * 6 and set(back) <= set(sd)               # integer bookkeeping tensors (num_batches_tracked) may be dropped
    assert all(np.array_equal(np.asarray(back[k]), sd[k].numpy()) for k in back)
    torch.save({"net": {"foo": torch.zeros(1)}}, str(tmp_path / "bad.pt"))
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "export_weights.py"), str(tmp_path / "bad.pt"), str(tmp_path / "x.ptdw")],
                       capture_output=True, text=True)
    assert r.returncode == 2 and "lacks" in r.stdout


def test_algorithmic_work_figures_of_the_bench():
    """bench.py's roofline uses SURVEY.md section 8d's algorithmic work: 2*9*Cin*Cout FLOP and 4*(Cin+Cout) bytes per output pixel with
    UNPADDED channel counts - 148 474 FLOP and 1 808 B per padded full-resolution pixel, 139.87 GFLOP / 1 703 MB at 736x1280."""
    import importlib.util
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("ptd_bench", os.path.join(root, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    t = bench.layer_table(736, 1280)
    assert len(t) == 28
    flops, nbytes = sum(f for f, _ in t.values()), sum(b for _, b in t.values())
    # bytes: the formula the survey states gives 1 697.05 MB; its per-layer table sums to 1 703.1 MB because the rows below 1/4
    # resolution carry a little more than (Cin + Cout) * 4 B per pixel - 0.35 % apart, the bench uses the formula
    assert abs(flops / 1e9 - 139.87) < 0.01 and abs(nbytes / 1e6 - 1697.05) < 0.01 and abs(nbytes / 1e6 / 1703.1 - 1) < 0.005
    px = 736 * 1280
    assert round(flops / px) == 148474 and abs(nbytes / px - 1808) < 8
    assert abs(t["enc1.l2a"][0] / 1e9 - 34.73) < 0.01 and abs(t["dec1.c1"][1] / 1e6 - 252.5) < 0.1      # rows of the section 8a table
    peaks = bench.measured_peaks()
    assert peaks["hbm"] > 1000 and peaks["bf16"] > 100
