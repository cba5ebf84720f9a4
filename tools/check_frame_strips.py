"""torchrun --nproc-per-node N tools/check_frame_strips.py [W H frames mode gated]
Multi-process parity check of ptd_frame_submit / ptd_frame_wait on row strips: every rank (a) renders the untiled frames on its own
GPU through the two reference call sites and (b) submits the same frames through its strip handles with two frames in flight while the
next is submitted, host buffers for the G-buffer and the denoised frame; its rows of both must be bit-identical.  gated = 1 creates the
path-tracer strips with PTD_PT_GATED_MAIL, i.e. path trace of frame k + 1 and denoiser of frame k on two streams.
Prints one line per rank; exit code 1 on any mismatch.  Test infrastructure, not product."""
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
frames = int(sys.argv[3]) if len(sys.argv) > 3 else 6
mode = sys.argv[4] if len(sys.argv) > 4 else "2xf16"
gated = (sys.argv[5] if len(sys.argv) > 5 else "1") == "1"
flags = {"tf32": capi.DN_TF32, "f16": capi.DN_F16, "2xf16": capi.DN_2XF16, "3xtf32": capi.DN_3XTF32}[mode]
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
wfile = os.path.join(tempfile.gettempdir(), "ptd_check_%d.ptdw" % rank)
weights.save_weights(weights.synthetic_state_dict(1234), wfile)
sc = capi.Scene(path=os.path.join(ROOT, "scenes", "hall_64x48.txt"))
sc.set_resolution(W, H)
full_pt, full_dn = capi.PathTracer(sc, device=local), capi.Denoiser(wfile, H, W, device=local, flags=flags)
cams = [capi.frame_camera(sc.camera[0], k) for k in range(frames)]
refs = []
for k, cam in enumerate(cams):
    g = full_pt.render_host(cam)
    refs.append((g, full_dn.forward_host(g, reset=(k == 0))))
pipe = tiling.StripPipeline(sc, wfile, rank, world, local, dist, dn_flags=flags, gated=gated)
r0, nr = pipe.pt_rows
hg = [torch.zeros(10, H, W, dtype=torch.float32).pin_memory() for _ in range(frames)]
hr = [torch.zeros(3, H, W, dtype=torch.float32).pin_memory() for _ in range(frames)]
bad = 0


def check(k):
    global bad
    ok_g = hg[k].numpy()[:, r0:r0 + nr].tobytes() == refs[k][0][:, r0:r0 + nr].tobytes()
    ok_y = hr[k].numpy()[:, r0:r0 + nr].tobytes() == refs[k][1][:, r0:r0 + nr].tobytes()
    outside = not hr[k].numpy()[:, :r0].any() and not hr[k].numpy()[:, r0 + nr:].any()        # nobody else's rows are touched
    if not (ok_g and ok_y and outside):
        bad += 1
        print("rank %d frame %d MISMATCH gbuf=%s rgb=%s outside-rows-untouched=%s" % (rank, k, ok_g, ok_y, outside), flush=True)


slots = capi.frame_slots()                                     # 3: two frames stay in flight behind the one being submitted
done = 0


def wait_and_check():
    global bad, done
    pipe.pt.frame_wait()
    k = done
    if k % 3 != 2:                                             # (every third frame was submitted without a G-buffer pointer)
        check(k)
    elif hr[k].numpy()[:, r0:r0 + nr].tobytes() != refs[k][1][:, r0:r0 + nr].tobytes():
        bad += 1
    done += 1


for k in range(frames):
    pipe.pt.frame_submit(pipe.dn, hr[k], hg[k] if k % 3 != 2 else None, cam=cams[k], reset=(k == 0))
    if k >= slots - 1:
        wait_and_check()
while done < frames:
    wait_and_check()
t = torch.tensor([bad], device="cuda")
dist.all_reduce(t)
print("rank %d/%d rows [%d,%d) mode %s %s: %d frame(s) %s" % (rank, world, r0, r0 + nr, mode, "two streams (gated mail)" if gated else "one stream", frames,
                                                              "BIT-EXACT vs untiled" if bad == 0 else "%d MISMATCHES" % bad), flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(1 if int(t.item()) else 0)
