"""TEST INFRASTRUCTURE ONLY - CPU restatement of the reference's recurrent denoising autoencoder forward
(hot path HP-2), training/recurrent_autoencoder_model.py:8-142, over plain torch fp32 CPU ops.

Semantics frozen by SURVEY.md decisions D1/D3: eval-mode BatchNorm (test.py:35), hidden state carried between
frames and zeroed when `reset` (forward(x, j) with j == 0, model.py:121-128), input zero-padded bottom/right
to a multiple of 32 and the output cropped.

Pinned by tests/test_oracle_dn.py against tests/golden/dn_*.npz, which tests/tools/make_golden_dn.py produced by
importing the reference's own AutoEncoder from /root/reference/training (same weights, same inputs).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
"""
import numpy as np
import torch
import torch.nn.functional as F

from ai_path_tracer_denoiser_b200.weights import conv_layers, BN_EPS, LRELU_SLOPE


def synthetic_gbuffer(H, W, seed=0, frame=0):
    """SURVEY.md section 8d synthetic input: ch0-2 U[0,1), ch3-5 unit vectors, ch6 U[0,20), ch7-9 U[0,1)."""
    rng = np.random.RandomState(seed * 1000003 + frame)
    x = np.empty((10, H, W), np.float32)
    x[0:3] = rng.uniform(0, 1, (3, H, W))
    n = rng.standard_normal((3, H, W))
    x[3:6] = n / np.sqrt((n * n).sum(0, keepdims=True))
    x[6] = rng.uniform(0, 20, (H, W))
    x[7:10] = rng.uniform(0, 1, (3, H, W))
    return x


class DenoiserOracle:
    def __init__(self, state_dict, dtype=torch.float32, threads=None, batch_stats=False):
        """batch_stats=True: BatchNorm in TRAINING mode (batch mean / biased batch variance of the current input) - what the module
        traced by training/convert_to_torchscript.py:26-30 computes, since the script never calls .eval()."""
        self.batch_stats = batch_stats
        if threads:
            torch.set_num_threads(threads)
        self.dtype = dtype
        self.sd = {k: torch.from_numpy(np.asarray(v)).to(dtype) for k, v in state_dict.items() if np.asarray(v).dtype.kind == "f"}
        self.layers = {name: (ck, bk, order) for name, ck, bk, _, _, order in conv_layers()}
        self.hidden = None

    def _cbl(self, x, name):
        ck, bk, order = self.layers[name]
        y = F.conv2d(x, self.sd[ck + ".weight"], self.sd[ck + ".bias"], padding=1)
        if self.batch_stats:
            bn = lambda t: F.batch_norm(t, None, None, self.sd[bk + ".weight"], self.sd[bk + ".bias"], training=True, eps=BN_EPS)
        else:
            bn = lambda t: F.batch_norm(t, self.sd[bk + ".running_mean"], self.sd[bk + ".running_var"], self.sd[bk + ".weight"],
                                        self.sd[bk + ".bias"], training=False, eps=BN_EPS)
        if order == "bn_lrelu":
            return F.leaky_relu(bn(y), LRELU_SLOPE)
        return bn(F.leaky_relu(y, LRELU_SLOPE))          # encoder layer2 first conv: LeakyReLU then BN (model.py:30-32)

    def _block(self, x, pre, i):
        out1 = self._cbl(x, pre + ".l1")                                         # model.py:66 / :76
        out2 = self._cbl(torch.cat((out1, self.hidden[i]), dim=1), pre + ".l2a")  # :67 / :77
        out2 = self._cbl(out2, pre + ".l2b")
        self.hidden[i] = out2                                                    # :68 / :79
        return out2

    def forward_padded(self, x, reset, taps=None):
        """x: torch [1,10,Hp,Wp], Hp,Wp multiples of 32.  taps: optional dict filled with per-layer activations."""
        _, _, H, W = x.shape
        assert H % 32 == 0 and W % 32 == 0
        if reset or self.hidden is None:                                         # model.py:121-128, :83-90
            ch = [32, 43, 57, 76, 101, 101]
            self.hidden = [torch.zeros(1, c, H >> i, W >> i, dtype=self.dtype) for i, c in enumerate(ch)]
        skips = []
        t = x
        for k in range(1, 6):                                                    # :129-133
            t = F.max_pool2d(self._block(t, "enc%d" % k, k - 1), 2)
            skips.append(t)
            if taps is not None:
                taps["e%d" % k] = t
        t = self._block(t, "bott", 5)                                            # :135
        if taps is not None:
            taps["b"] = t
        for k in (5, 4, 3, 2, 1):                                                # :136-140
            t = torch.cat((t, skips[k - 1]), dim=1)
            t = F.interpolate(t, scale_factor=2, mode="nearest")                 # model.py:40
            t = self._cbl(t, "dec%d.c1" % k)
            t = self._cbl(t, "dec%d.c2" % k)
            if taps is not None:
                taps["d%d" % k] = t
        return t

    def forward(self, gbuf, reset):
        """gbuf: numpy [10,H,W] fp32 (any H,W) -> numpy [3,H,W]; pad/crop policy D3."""
        _, H, W = gbuf.shape
        Hp, Wp = (H + 31) // 32 * 32, (W + 31) // 32 * 32
        x = torch.zeros(1, 10, Hp, Wp, dtype=self.dtype)
        x[0, :, :H, :W] = torch.from_numpy(np.ascontiguousarray(gbuf)).to(self.dtype)
        with torch.no_grad():
            y = self.forward_padded(x, reset)
        return y[0, :, :H, :W].to(torch.float32).numpy().copy()
