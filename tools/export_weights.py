"""python tools/export_weights.py CHECKPOINT.pt OUT.ptdw      - training checkpoint -> the flat PTDW file ptd_dn_create reads
   python tools/export_weights.py --synthetic 1234 OUT.ptdw   - the seeded synthetic weights the tests / bench use (no checkpoint ships)

Replaces the reference's TorchScript export (training/convert_to_torchscript.py:26-30): it loads the same `{'net': state_dict}`
checkpoints train.py writes (train.py:108-112) and dumps the 196 state-dict tensors; BatchNorm is applied in eval mode by the
denoiser (running statistics), hidden state is carried by the caller's `reset_hidden` flag."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ai_path_tracer_denoiser_b200 import weights  # noqa: E402


def main(argv):
    if len(argv) == 3 and argv[0] == "--synthetic":
        sd = weights.synthetic_state_dict(int(argv[1]))
        out = argv[2]
    elif len(argv) == 2:
        import torch
        ck = torch.load(argv[0], map_location="cpu", weights_only=False)
        sd = ck["net"] if isinstance(ck, dict) and "net" in ck else ck
        if hasattr(sd, "state_dict"):
            sd = sd.state_dict()
        out = argv[1]
    else:
        print(__doc__)
        return 1
    need = []
    for name, conv_key, bn_key, ci, co, _ in weights.conv_layers():
        need += [conv_key + ".weight", conv_key + ".bias", bn_key + ".weight", bn_key + ".bias", bn_key + ".running_mean", bn_key + ".running_var"]
    missing = [k for k in need if k not in sd]
    if missing:
        print("checkpoint lacks %d tensors of recurrent_autoencoder_model.AutoEncoder, e.g. %s" % (len(missing), missing[:3]))
        return 2
    weights.save_weights(sd, out)
    print("wrote %s (%d tensors)" % (out, len(sd)))
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv[1:]))
