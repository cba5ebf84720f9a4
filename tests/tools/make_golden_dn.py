"""Generates tests/golden/dn_*.npz by running the REFERENCE's own model definition
(/root/reference/training/recurrent_autoencoder_model.py, imported, not copied) on CPU in eval mode with
the synthetic seeded weights of ai_path_tracer_denoiser_b200.weights, over a short frame sequence
(j = 0, 1, 2 -> hidden state carried, as training/test.py:42-48 does).  Run in the build container only."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference/training")
import recurrent_autoencoder_model as ref_model  # noqa: E402  (the reference file itself)

from ai_path_tracer_denoiser_b200.weights import synthetic_state_dict  # noqa: E402
from oracle.dn_oracle import synthetic_gbuffer  # noqa: E402


def main():
    torch.manual_seed(0)
    torch.set_num_threads(8)
    sd = synthetic_state_dict(1234)
    model = ref_model.AutoEncoder(10)
    missing = model.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()}, strict=True)
    print("load_state_dict:", missing, "tensors:", len(sd), "params:", sum(p.numel() for p in model.parameters()))
    model.eval()
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    for (H, W, frames) in ((64, 96, 3), (32, 32, 2)):
        xs = np.stack([synthetic_gbuffer(H, W, seed=5, frame=j) for j in range(frames)])
        ys = []
        with torch.no_grad():
            for j in range(frames):
                ys.append(model(torch.from_numpy(xs[j:j + 1]), j)[0].numpy().copy())
        ys = np.stack(ys)
        hid = model.encoder3[0].hidden[0].numpy().copy()
        np.savez_compressed(os.path.join(out_dir, "dn_%dx%d.npz" % (H, W)), x=xs.astype(np.float16) if False else xs, y=ys, hidden3=hid)
        print(H, W, "out abs-max", np.abs(ys).max(axis=(1, 2, 3)))
    # what the TorchScript export of the reference actually computes - convert_to_torchscript.py:26-30 traces
    # model.forward (j defaults to 0: hidden state zeroed every call) without ever calling .eval(): BatchNorm in training mode
    model.train()
    H, W, frames = 96, 160, 2
    xs = np.stack([synthetic_gbuffer(H, W, seed=9, frame=j) for j in range(frames)])
    with torch.no_grad():
        ys = np.stack([model(torch.from_numpy(xs[j:j + 1]))[0].numpy().copy() for j in range(frames)])
    np.savez_compressed(os.path.join(out_dir, "dn_traced_%dx%d.npz" % (H, W)), x=xs, y=ys)
    print("traced-export semantics", H, W, "out abs-max", np.abs(ys).max(axis=(1, 2, 3)))


if __name__ == "__main__":
    main()
