"""Denoiser weight container: naming, synthetic initialisation and the flat file the C-ABI loads.

The reference keeps its weights in a PyTorch checkpoint `{'net': state_dict}` (training/train.py:108-112)
that reaches the renderer as a TorchScript file (Inference/src/main.cpp:39,107).  No checkpoint ships with
the reference, so (SURVEY.md decision D2) the parity and bench weights are synthetic and seeded:
conv = kaiming-normal fan_in, bias 0.01 (exactly training/train.py:32-38); BatchNorm gamma~U(0.4,1.0),
beta,mean~N(0,0.1), var~U(0.5,1.5) so that BN is non-trivial and the recurrence stays contractive.

The flat file ("PTDW") is what ptd_dn_create() reads: it is a plain dump of the state_dict, so a real
checkpoint can be exported with tools/export_weights.py the same way.
"""
import struct
from collections import OrderedDict

import numpy as np

ENC = [(10, 32), (32, 43), (43, 57), (57, 76), (76, 101)]     # recurrent_autoencoder_model.py:98-107
BOTT = (101, 101)                                             # :109
DEC = [(101, 76), (76, 57), (57, 43), (43, 32), (32, 3)]      # :111-115 (decoder5 .. decoder1)
BN_EPS = 1e-5
LRELU_SLOPE = 0.1


def conv_layers():
    """The 28 convolutions in execution order: (name, conv key, bn key, cin, cout, epilogue order)."""
    L = []
    for k, (ci, co) in enumerate(ENC, 1):
        p = "encoder%d.0." % k
        L.append(("enc%d.l1" % k, p + "layer1.0", p + "layer1.1", ci, co, "bn_lrelu"))      # :23-27
        L.append(("enc%d.l2a" % k, p + "layer2.0", p + "layer2.2", 2 * co, co, "lrelu_bn"))  # :30-32
        L.append(("enc%d.l2b" % k, p + "layer2.3", p + "layer2.4", co, co, "bn_lrelu"))      # :33-35
    ci, co = BOTT
    L.append(("bott.l1", "bottleneck.layer1.0", "bottleneck.layer1.1", ci, co, "bn_lrelu"))  # :50-54
    L.append(("bott.l2a", "bottleneck.layer2.0", "bottleneck.layer2.1", 2 * co, co, "bn_lrelu"))
    L.append(("bott.l2b", "bottleneck.layer2.3", "bottleneck.layer2.4", co, co, "bn_lrelu"))
    for k, (ci, co) in zip((5, 4, 3, 2, 1), DEC):
        p = "decoder%d.layer1." % k
        L.append(("dec%d.c1" % k, p + "1", p + "2", 2 * ci, co, "bn_lrelu"))                 # :40-43
        L.append(("dec%d.c2" % k, p + "4", p + "5", co, co, "bn_lrelu"))                     # :44-46
    return L


def synthetic_state_dict(seed=1234):
    """OrderedDict name -> float32 ndarray with the reference's state_dict keys (196 tensors)."""
    rng = np.random.RandomState(seed)
    sd = OrderedDict()
    for _, ck, bk, ci, co, _ in conv_layers():
        std = np.sqrt(2.0 / (ci * 9))
        sd[ck + ".weight"] = (rng.standard_normal((co, ci, 3, 3)) * std).astype(np.float32)
        sd[ck + ".bias"] = np.full((co,), 0.01, np.float32)
        sd[bk + ".weight"] = rng.uniform(0.4, 1.0, co).astype(np.float32)
        sd[bk + ".bias"] = (rng.standard_normal(co) * 0.1).astype(np.float32)
        sd[bk + ".running_mean"] = (rng.standard_normal(co) * 0.1).astype(np.float32)
        sd[bk + ".running_var"] = rng.uniform(0.5, 1.5, co).astype(np.float32)
        sd[bk + ".num_batches_tracked"] = np.zeros((), np.int64)
    return sd


def save_weights(sd, path):
    """PTDW v1: magic, u32 version, u32 count; per tensor: u32 name_len, name, u32 ndim, u32 dims[], f32 data.
    Integer buffers (num_batches_tracked) are skipped - inference never reads them."""
    items = [(k, np.asarray(v)) for k, v in sd.items() if np.asarray(v).dtype.kind == "f"]
    with open(path, "wb") as f:
        f.write(b"PTDW" + struct.pack("<II", 1, len(items)))
        for k, v in items:
            kb = k.encode()
            v = np.ascontiguousarray(v, np.float32)
            f.write(struct.pack("<I", len(kb)) + kb + struct.pack("<I", v.ndim) + struct.pack("<%dI" % v.ndim, *v.shape))
            f.write(v.tobytes())
    return path


def load_weights(path):
    sd = OrderedDict()
    with open(path, "rb") as f:
        assert f.read(4) == b"PTDW"
        ver, n = struct.unpack("<II", f.read(8))
        assert ver == 1
        for _ in range(n):
            (kl,) = struct.unpack("<I", f.read(4))
            k = f.read(kl).decode()
            (nd,) = struct.unpack("<I", f.read(4))
            shape = struct.unpack("<%dI" % nd, f.read(4 * nd))
            cnt = int(np.prod(shape)) if nd else 1
            sd[k] = np.frombuffer(f.read(4 * cnt), np.float32).reshape(shape).copy()
    return sd
