"""Host-side logic of the multi-GPU row-strip tiling (ai_path_tracer_denoiser_b200/tiling.py) on CPU: two processes over the
`gloo` backend exchange their strips' blobs, agree on the partition and pick their neighbours.  No GPU, no compute call."""
import os
import socket
import sys

import pytest
import torch.multiprocessing as mp

from conftest import ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, H, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from ai_path_tracer_denoiser_b200 import capi, tiling
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        dn_rows, pt_rows = tiling.strip_rows(H, world, rank)
        blob = b"strip-%d:" % rank + bytes(dn_rows) if max(dn_rows) < 256 else b"strip-%d:%d,%d" % (rank, dn_rows[0], dn_rows[1])
        blob = blob.ljust(64, b".")                                   # equal-sized PODs, like ptd_dn_strip_export's
        blobs = tiling.exchange_blobs(blob, dist, world)
        up, down = tiling.neighbours(blobs, rank)
        # the real constructors must fail loudly here (no CUDA device): nothing falls back to the CPU
        loud = False
        if capi.device_count() == 0:
            try:
                sc = capi.Scene(path=os.path.join(ROOT, "scenes", "cornell_64x48.txt"))
                tiling.StripPipeline(sc, "/nonexistent.ptdw", rank, world, 0, dist)
            except capi.PtdError as e:
                loud = "no CPU fallback" in str(e)
        q.put((rank, dn_rows, pt_rows, [b[:8] for b in blobs], up[:8] if up else None, down[:8] if down else None, loud))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("H", [720, 1080])
def test_two_ranks_agree_on_strips_and_neighbours(H):
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, H, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, dn0, pt0, blobs0, up0, down0, loud0), (r1, dn1, pt1, blobs1, up1, down1, loud1) = res
    Hp = (H + 31) // 32 * 32
    assert dn0[0] == 0 and dn0[0] + dn0[1] == dn1[0] and dn1[0] + dn1[1] == Hp and dn0[1] % 32 == 0 and dn1[1] % 32 == 0
    assert pt0 == (0, dn0[1]) and pt1 == (dn1[0], H - dn1[0])        # the path tracer's strips stop at the real frame height
    assert blobs0 == blobs1 == [b"strip-0:", b"strip-1:"]
    assert up0 is None and down0 == b"strip-1:" and up1 == b"strip-0:" and down1 is None
    import torch
    if not torch.cuda.is_available():
        assert loud0 and loud1
