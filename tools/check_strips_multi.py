"""torchrun --nproc-per-node N tools/check_strips_multi.py [W H frames]
Multi-process parity check of the row-strip mode: every rank runs (a) the untiled frame loop on its own GPU and (b) its strip
of the tiled loop (IPC-connected to the other ranks), and compares its rows of the denoised frame and of the G-buffer bit for
bit.  Prints one line per rank; exit code 1 on any mismatch.  Test infrastructure, not product."""
import ctypes as C
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
from ai_path_tracer_denoiser_b200 import capi, tiling, weights  # noqa: E402

W = int(sys.argv[1]) if len(sys.argv) > 1 else 320
H = int(sys.argv[2]) if len(sys.argv) > 2 else 250
frames = int(sys.argv[3]) if len(sys.argv) > 3 else 4
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
wfile = os.path.join(tempfile.gettempdir(), "ptd_check_%d.ptdw" % rank)
weights.save_weights(weights.synthetic_state_dict(1234), wfile)
sc = capi.Scene(path=os.path.join(ROOT, "scenes", "hall_64x48.txt"))
sc.set_resolution(W, H)
full_pt, full_dn = capi.PathTracer(sc, device=local), capi.Denoiser(wfile, H, W, device=local, flags=capi.DN_TF32)
pipe = tiling.StripPipeline(sc, wfile, rank, world, local, dist)
P = W * H
g = torch.zeros(10 * P, dtype=torch.float32, device="cuda")
out = torch.zeros(3 * P, dtype=torch.float32, device="cuda")
r0, nr = pipe.pt_rows
bad = 0
for k in range(frames):
    cam = capi.frame_camera(sc.camera[0], k)
    ref_g = full_pt.render_host(cam)
    ref = full_dn.forward_host(ref_g, reset=(k == 0))
    pipe.frame(cam, C.c_void_p(g.data_ptr()), C.c_void_p(out.data_ptr()), k == 0)
    torch.cuda.synchronize()
    y = out.cpu().numpy().reshape(3, H, W)[:, r0:r0 + nr]
    gg = g.cpu().numpy().reshape(10, H, W)[:, r0:r0 + nr]
    ok_g = gg.tobytes() == ref_g[:, r0:r0 + nr].tobytes()
    ok_y = y.tobytes() == ref[:, r0:r0 + nr].tobytes()
    if not (ok_g and ok_y):
        bad += 1
        print("rank %d frame %d MISMATCH gbuf=%s rgb=%s maxdiff=%g" % (rank, k, ok_g, ok_y, float(np.abs(y - ref[:, r0:r0 + nr]).max())), flush=True)
t = torch.tensor([bad], device="cuda")
dist.all_reduce(t)
print("rank %d/%d rows [%d,%d): %d frame(s) %s" % (rank, world, r0, r0 + nr, frames, "BIT-EXACT vs untiled" if bad == 0 else "%d MISMATCHES" % bad), flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(1 if int(t.item()) else 0)
