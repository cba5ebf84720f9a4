"""Row-strip tiling of one frame over the GPUs of a box: one process per GPU (torchrun), torch.distributed only for the
host-side plumbing (exchanging the strips' IPC blobs once at start-up, barriers); the data path is libptd.so's kernels
storing halo rows / live counts straight into the neighbours' memory over NVLink (include/ptd.h, "row-strip mode").

    rank r owns padded rows [row0, row0 + rows) of the frame (32-row groups split evenly, SURVEY.md 8e / decision D7):
      path tracer strip   image rows  [row0, min(row0 + rows, H))
      denoiser strip      padded rows [row0, row0 + rows)
"""
import os

import numpy as np

from . import capi


def strip_rows(H, world, rank):
    """((dn_row0, dn_rows), (pt_row0, pt_rows)) of `rank`."""
    r0, rows = capi.strip_partition(H, world, rank)
    return (r0, rows), (r0, min(r0 + rows, H) - r0)


def exchange_blobs(blob, dist=None, world=1):
    """All-gather one bytes blob per rank (host side, any backend).  Returns the list in rank order."""
    if dist is None or world == 1:
        return [blob]
    import torch
    t = torch.frombuffer(bytearray(blob), dtype=torch.uint8).clone()
    dev = None
    if dist.get_backend() == "nccl":
        dev = torch.device("cuda", torch.cuda.current_device())
        t = t.to(dev)
    out = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(out, t)
    return [bytes(o.cpu().numpy().tobytes()) for o in out]


def neighbours(blobs, rank):
    """(up, down) blobs of `rank` (None at the frame border)."""
    return (blobs[rank - 1] if rank > 0 else None), (blobs[rank + 1] if rank + 1 < len(blobs) else None)


class StripPipeline:
    """Path tracer + denoiser of this rank's strip, connected to the other ranks' strips."""

    def __init__(self, scene, weights_path, rank, world, device, dist=None, dn_flags=capi.DN_TF32, gated=None):
        cam = scene.camera[0]
        self.W, self.H = int(cam["res"][0]), int(cam["res"][1])
        self.rank, self.world = rank, world
        self.dn_rows, self.pt_rows = strip_rows(self.H, world, rank)
        if world == 1:
            self.pt = capi.PathTracer(scene, device=device)
            self.dn = capi.Denoiser(weights_path, self.H, self.W, device=device, flags=dn_flags)
            return
        # PTD_STRIP_PIPELINE=1 (same value on every rank): the two-stream frame loop in strip mode.  It needs the gated live-count
        # mail - with it no path-trace block ever spins on another GPU, which is what made the two-stream loop deadlock (FrameLoop)
        self.two_stream_ok = os.environ.get("PTD_STRIP_PIPELINE", "0") == "1" if gated is None else bool(gated)
        self.pt = capi.PathTracer(scene, device=device, strip=self.pt_rows, flags=capi.PT_GATED_MAIL if self.two_stream_ok else 0)
        self.dn = capi.Denoiser(weights_path, self.H, self.W, device=device, flags=dn_flags, strip=self.dn_rows)
        pt_blobs = exchange_blobs(self.pt.export_info(), dist, world)
        dn_blobs = exchange_blobs(self.dn.export_info(), dist, world)
        self.pt.connect(pt_blobs, rank)
        self.dn.connect(dn_blobs, rank)
        if dist is not None:
            dist.barrier()

    def frame(self, cam, gbuf_ptr, rgb_ptr, reset, stream=None):
        """One frame of the render loop for this strip (asynchronous on `stream`)."""
        self.pt.render(gbuf_ptr, cam=np.ascontiguousarray(cam), iter=1, stream=stream)
        self.dn.forward(gbuf_ptr, rgb_ptr, reset, stream=stream)


class FrameLoop:
    """The render loop (the reference's runCuda(), main.cpp:120-168) on top of a StripPipeline: frame k = path trace with camera k,
    then the denoiser with the hidden state of frame k - 1.

    pipelined=True runs the two hot paths on two CUDA streams with a double-buffered G-buffer: the path trace of frame k + 1 (which
    depends on nothing but its camera) overlaps the denoiser of frame k.  Per-frame latency is unchanged, throughput rises wherever
    one of the two leaves the GPU idle - above all in the multi-GPU strip mode, where both are chains of short, latency-bound
    kernels.  Frames are still produced in order and the recurrent state is carried exactly as in the serial loop (the denoiser
    stream is sequential), so the output is bit-identical to the serial loop's."""

    def __init__(self, pipe, pipelined=True):
        import ctypes
        import torch
        self._C, self._torch = ctypes, torch
        # Strip mode stays SERIAL.  With two streams a denoiser conv of frame k that spins on a neighbour's flag holds ~200 KB of
        # every SM's shared memory, so the path-trace shade blocks of frame k + 1 (41 KB each) cannot become resident beside it; if
        # the neighbour's SMs are in turn held by ITS frame-k+1 shade blocks spinning on OUR frame-k+1 live count, nobody moves.
        # Seen once as a hang at N = 4 (measured before that: 562 / 798 frames/s at N = 4 / 8 instead of 468 / 597).  On one
        # stream every wait points at a kernel that is earlier in every rank's order, so the serial loop cannot deadlock.
        # PTD_STRIP_PIPELINE=1 re-enables it on top of PTD_PT_GATED_MAIL: the mail wait moves into a one-warp gate kernel, shade blocks
        # never spin, so a neighbour's convs can always become resident and every conv's wait points at a kernel that will run
        # (opt-in until it has been soaked on 4 and 8 GPUs; every device-side wait now traps after 20 s instead of hanging).
        if pipe.world > 1 and not getattr(pipe, "two_stream_ok", False):
            pipelined = False
        self.pipe, self.pipelined = pipe, pipelined
        P = pipe.W * pipe.H
        self.s_pt = torch.cuda.Stream()
        # (a higher priority for the denoiser stream was measured to make no difference: 222.0 vs 222.9 frames/s)
        prio = int(os.environ.get("PTD_DN_STREAM_PRIORITY", "0"))
        self.s_dn = torch.cuda.Stream(priority=prio) if pipelined else self.s_pt
        n = 2 if pipelined else 1
        self.gbuf = [torch.zeros(10 * P, dtype=torch.float32, device="cuda") for _ in range(n)]
        self.rgb = torch.zeros(3 * P, dtype=torch.float32, device="cuda")
        self.ev_pt = [torch.cuda.Event() for _ in range(n)]
        self.ev_dn = [None] * n
        self.k = 0

    def frame(self, cam, reset):
        """Enqueue one frame (asynchronous).  Returns the index of the G-buffer it uses."""
        C, torch = self._C, self._torch
        i = self.k % len(self.gbuf)
        g = C.c_void_p(self.gbuf[i].data_ptr())
        if self.pipelined and self.ev_dn[i] is not None:
            self.s_pt.wait_event(self.ev_dn[i])                 # the denoiser of frame k - 2 has consumed this G-buffer
        self.pipe.pt.render(g, cam=np.ascontiguousarray(cam), iter=1, stream=C.c_void_p(self.s_pt.cuda_stream))
        if self.pipelined:
            self.ev_pt[i].record(self.s_pt)
            self.s_dn.wait_event(self.ev_pt[i])
        self.pipe.dn.forward(g, C.c_void_p(self.rgb.data_ptr()), reset, stream=C.c_void_p(self.s_dn.cuda_stream))
        if self.pipelined:
            if self.ev_dn[i] is None:
                self.ev_dn[i] = torch.cuda.Event()
            self.ev_dn[i].record(self.s_dn)
        self.k += 1
        return i

    def synchronize(self):
        self.s_pt.synchronize()
        self.s_dn.synchronize()
