"""GPU-box report: max-abs / rel-L2 error of every denoiser mode against the torch-fp32-CPU oracle over a 12-frame recurrence
(no reset after frame 0) at 160x224, outputs O(1).  Test infrastructure: writes the table DESIGN.md quotes."""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from ai_path_tracer_denoiser_b200 import capi, weights  # noqa: E402
from oracle.dn_oracle import DenoiserOracle, synthetic_gbuffer  # noqa: E402

H, W, frames = 160, 224, 12
sd = weights.synthetic_state_dict(1234)
wfile = weights.save_weights(sd, os.path.join(tempfile.gettempdir(), "ptd_acc.ptdw"))
O = DenoiserOracle(sd)
refs = []
for j in range(frames):
    refs.append(O.forward(synthetic_gbuffer(H, W, seed=21, frame=j), reset=(j == 0)))
print("mode      frame0 max-abs  rel-L2     last-frame max-abs  rel-L2     (|out| max %.2f)" % max(float(np.abs(r).max()) for r in refs))
for name, flag in (("fp32", capi.DN_FP32), ("3xtf32", capi.DN_3XTF32), ("tf32", capi.DN_TF32), ("f16", capi.DN_F16)):
    dn = capi.Denoiser(wfile, H, W, flags=flag)
    errs = []
    for j in range(frames):
        y = dn.forward_host(synthetic_gbuffer(H, W, seed=21, frame=j), reset=(j == 0))
        d = y.astype(np.float64) - refs[j]
        errs.append((np.abs(d).max(), np.sqrt((d * d).sum() / (refs[j].astype(np.float64) ** 2).sum())))
    print("%-8s  %.3e       %.3e  %.3e           %.3e" % (name, errs[0][0], errs[0][1], errs[-1][0], errs[-1][1]))
