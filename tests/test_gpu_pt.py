"""GPU parity tests of the path-trace hot path (HP-1), through the C ABI (libptd.so).

Checker hierarchy:
  * oracle A  = the UNMODIFIED reference kernels (oracle/_ref/libref_pt.so: pathtrace.cu compiled verbatim for sm_100a,
                prebuilt in the build container, travels to the GPU box) run on the same GPU: everything bit-exact -
                PathSegment arrays entering every bounce (i.e. after every compaction / sort), ShadeableIntersections,
                the final thrust::partition layout, the image and the 10-plane G-buffer.
  * golden    = tests/golden/pt_*.npz from the reference's host build (no FMA, glibc libm): the integer fields
                (pixelIndex, remainingBounces) must agree up to a handful of borderline rays (reported, bounded).
  * oracle    = oracle/pt_oracle.c, for sizes the CPU finishes in seconds.
"""
import glob
import os

import numpy as np
import pytest

from conftest import GOLDEN, SCENES

pytestmark = pytest.mark.gpu

GOLD = sorted(glob.glob(os.path.join(GOLDEN, "pt_*.npz")))


def _capi():
    from ai_path_tracer_denoiser_b200 import capi
    if capi.device_count() < 1:
        pytest.fail("no CUDA device visible to libptd.so - the product has no CPU fallback")
    return capi


def _same(a, b, what, skip=("pad",)):
    assert len(a) == len(b), "%s: length %d vs %d" % (what, len(a), len(b))
    for f in a.dtype.names:
        if f in skip:
            continue
        if a[f].tobytes() != b[f].tobytes():
            bad = np.nonzero(np.any(np.atleast_2d((a[f].view(np.uint32) if a[f].dtype != np.uint8 else a[f]).reshape(len(a), -1) !=
                                                  (b[f].view(np.uint32) if b[f].dtype != np.uint8 else b[f]).reshape(len(b), -1)), axis=1))[0]
            raise AssertionError("%s field %s: %d of %d records differ, first at %d: %r vs %r" % (what, f, len(bad), len(a), bad[0], a[f][bad[0]], b[f][bad[0]]))


def _render_ours(capi, arrays, cam, flags):
    a = dict(arrays)
    a["camera"] = cam
    sc = capi.Scene(arrays=a)
    pt = capi.PathTracer(sc, flags=flags | capi.PT_TRACE | capi.PT_KEEP_TERMINATED)
    tensor = pt.render_host()
    counts, run = pt.live_counts()
    trace = [dict(paths=pt.dump_paths(b), isx=pt.dump_intersections(b)) for b in range(run)]
    # Without the reject array (the product configuration) pt_shade stores a tile's survivors one pipeline step later: same frame, bit for bit.
    for fl in (flags | capi.PT_TRACE, flags):
        pt2 = capi.PathTracer(sc, flags=fl)
        assert pt2.render_host().tobytes() == tensor.tobytes()
        assert pt2.live_counts()[0][:run] == counts[:run]
        assert pt2.dump_image().tobytes() == pt.dump_image().tobytes()
        if fl & capi.PT_TRACE:
            for b in range(run):
                _same(pt2.dump_paths(b), trace[b]["paths"], "deferred stores, bounce %d paths" % b)
                _same(pt2.dump_intersections(b), trace[b]["isx"], "deferred stores, bounce %d intersections" % b)
    return dict(tensor=tensor, counts=counts[:run], trace=trace, final=pt.dump_final_paths(), image=pt.dump_image(), pt=pt)


def _load_gold(path):
    g = np.load(path)
    arrays = dict(geoms=g["geoms"], materials=g["materials"], faces=g["faces"], mesh_box=g["mesh_box"], depth=int(g["depth"]), iterations=1)
    return g, arrays


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[3:-4] for p in GOLD])
def test_bit_exact_vs_reference_kernels(path):
    """Ours vs oracle A on the same GPU: bit-exact everything."""
    capi = _capi()
    from oracle import reflib
    g, arrays = _load_gold(path)
    sort = bool(g["sort"])
    variant = "sort" if sort else ""
    if not reflib.available(variant):
        pytest.fail("oracle/_ref/libref_pt%s.so missing on the GPU box (build it with `make -C oracle ref` before gpurun)" % ("_" + variant if variant else ""))
    R = reflib.RefLib(variant)
    name = os.path.basename(path)[3:-4]
    scene_file = os.path.join(SCENES, name.split("_f")[0] + ".txt")
    s = R.load_scene(scene_file)
    R.set_camera(s, g["camera"])
    ref = R.gpu_render(s, trace=True)
    ours = _render_ours(capi, arrays, g["camera"], capi.PT_SORT_MATERIAL if sort else 0)
    assert [b["n"] for b in ref["trace"]] == ours["counts"]
    for b, (r, o) in enumerate(zip(ref["trace"], ours["trace"])):
        _same(o["paths"], r["paths"], "bounce %d paths" % b)
        _same(o["isx"], r["isx"], "bounce %d intersections" % b)
    _same(ours["final"], ref["final_paths"], "final paths")
    assert ours["image"].tobytes() == ref["image"].tobytes()
    assert ours["tensor"].tobytes() == ref["tensor"].tobytes()


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[3:-4] for p in GOLD])
def test_indices_vs_cpu_golden(path):
    """Ours vs the committed golden vectors of the reference's HOST build (different libm / no FMA): the compacted
    (pixelIndex, remainingBounces) sequences may differ only through borderline rays; bound: <= 0.5 % of the records
    of any bounce, G-buffer planes 3-9 (first hit) equal within 1e-4 on >= 99.9 % of the pixels."""
    capi = _capi()
    g, arrays = _load_gold(path)
    ours = _render_ours(capi, arrays, g["camera"], capi.PT_SORT_MATERIAL if bool(g["sort"]) else 0)
    gold_counts = list(g["counts"])
    assert len(ours["counts"]) == len(gold_counts)
    for b, (n_o, n_g) in enumerate(zip(ours["counts"], gold_counts)):
        assert abs(n_o - n_g) <= max(2, 0.005 * n_g), "bounce %d live count %d vs %d" % (b, n_o, n_g)
    assert np.array_equal(ours["trace"][0]["paths"]["pix"], g["paths0"]["pix"])
    if ours["counts"][1] == gold_counts[1]:
        diff = np.count_nonzero(ours["trace"][1]["paths"]["pix"] != g["paths1"]["pix"])
        assert diff <= 0.005 * gold_counts[1]
    close = np.isclose(ours["tensor"][3:], g["tensor"][3:], rtol=0, atol=1e-4)
    assert close.mean() >= 0.999


def test_bvh_equals_brute_force():
    """BVH traversal must reproduce the reference's brute-force nearest hit incl. tie-breaks: bit-exact against PTD_PT_NO_BVH."""
    capi = _capi()
    for name in ("pt_hall_64x48_f0.npz", "pt_hall_reflective_64x48_f150.npz"):
        g, arrays = _load_gold(os.path.join(GOLDEN, name))
        a = _render_ours(capi, arrays, g["camera"], 0)
        b = _render_ours(capi, arrays, g["camera"], capi.PT_NO_BVH)
        st = a["pt"].bvh_stats()
        assert st["nodes"] > 1 and st["max_leaf"] <= 16
        assert a["counts"] == b["counts"]
        for x, y in zip(a["trace"], b["trace"]):
            _same(x["paths"], y["paths"], "paths")
            _same(x["isx"], y["isx"], "isx")
        assert a["tensor"].tobytes() == b["tensor"].tobytes()


def test_scene_file_render_matches_oracle_larger():
    """Scene-file path (our parser + OBJ reader + BVH) at a size the C oracle finishes in seconds, and properties that
    hold at any size: compaction conserves paths (every pixel terminates exactly once), counts are non-increasing,
    every live pixelIndex is unique."""
    capi = _capi()
    from oracle import pt_oracle
    sc = capi.Scene(path=os.path.join(SCENES, "hall_64x48.txt"))
    sc.set_resolution(160, 96)
    cam = capi.frame_camera(sc.camera[0], 17)
    sc.set_camera(cam)
    A = sc.arrays()
    pt = capi.PathTracer(sc, flags=capi.PT_TRACE | capi.PT_KEEP_TERMINATED)
    tensor = pt.render_host()
    counts, run = pt.live_counts()
    ora = pt_oracle.render(A, cam, trace=True)
    for b in range(run):
        assert abs(counts[b] - ora["counts"][b]) <= max(2, 0.005 * ora["counts"][b])
        pix = pt.dump_paths(b)["pix"]
        assert len(np.unique(pix)) == len(pix)
        assert np.all(np.diff(counts[:run]) <= 0)
    final = pt.dump_final_paths()
    assert np.array_equal(np.sort(final["pix"]), np.arange(160 * 96))
    assert np.all(final["rb"] == 0)
    close = np.isclose(tensor[3:], ora["tensor"][3:], rtol=0, atol=1e-4)
    assert close.mean() >= 0.999


def test_full_size_properties_720p():
    """BASELINE config sizes (1280x720 Cornell): size-independent properties + KAT live counts of SURVEY.md section 8d."""
    capi = _capi()
    from ai_path_tracer_denoiser_b200 import scenegen
    import tempfile
    d = tempfile.mkdtemp()
    path = scenegen.write_cornell(os.path.join(d, "c2.txt"), 1280, 720)
    sc = capi.Scene(path=path)
    sc.set_camera(capi.frame_camera(sc.camera[0], 0))
    pt = capi.PathTracer(sc, flags=capi.PT_KEEP_TERMINATED)
    t1 = pt.render_host()
    counts, run = pt.live_counts()
    assert counts[0] == 921600 and run == 8
    survey = [921600, 423505, 294406, 231943, 189033, 155133, 126972, 104657]
    for a, b in zip(counts, survey):
        assert abs(a - b) <= 0.002 * b
    final = pt.dump_final_paths()
    assert np.array_equal(np.sort(final["pix"]), np.arange(921600))
    t2 = pt.render_host()                                   # fixed seed: every frame identical (main.cpp:164 camchanged)
    assert t1.tobytes() == t2.tobytes()
    assert np.isfinite(t1).all()


def pt_strip_rows(capi, H, n):
    """Image rows of the n strips: the denoiser's 32-row-aligned partition of the padded frame, clipped to H."""
    parts = [capi.strip_partition(H, n, i) for i in range(n)]
    return [(r0, min(r0 + rows, H) - r0) for r0, rows in parts]


@pytest.mark.parametrize("scene,W,H,n", [("hall_64x48.txt", 96, 80, 3), ("cornell_specular_64x48.txt", 64, 72, 2)])
def test_row_strips_equal_the_full_frame(scene, W, H, n):
    """Multi-GPU tiling of the path tracer on ONE device: row strips, each its own handle; the frame-wide compacted index of the
    RNG seed is rebuilt from the live counts the strips mail each other.  Everything must be bit-identical to the untiled
    render: the G-buffer, the live counts (summed) and the concatenated per-bounce PathSegment arrays."""
    capi = _capi()
    import torch
    sc = capi.Scene(path=os.path.join(SCENES, scene))
    sc.set_resolution(W, H)
    cam = capi.frame_camera(sc.camera[0], 5)
    sc.set_camera(cam)
    full = capi.PathTracer(sc, flags=capi.PT_TRACE)
    ref = full.render_host()
    ref_counts, run = full.live_counts()
    rows = pt_strip_rows(capi, H, n)
    assert sum(r for _, r in rows) == H
    strips = [capi.PathTracer(sc, flags=capi.PT_TRACE, strip=r) for r in rows]
    infos = [s.export_info() for s in strips]
    for i, s in enumerate(strips):
        s.connect(infos, i)
    g = torch.zeros(10 * H * W, dtype=torch.float32, device="cuda")
    for rep in range(2):                                           # second frame: the mailboxes carry a new epoch
        g.zero_()
        capi.PathTracer.render_group(strips, [g.data_ptr()] * n)
        torch.cuda.synchronize()
        assert g.cpu().numpy().reshape(10, H, W).tobytes() == ref.tobytes()
    counts = [s.live_counts()[0] for s in strips]
    assert [sum(c[b] for c in counts) for b in range(run)] == ref_counts[:run]
    for b in range(run):
        cat = np.concatenate([s.dump_paths(b) for s in strips])
        _same(cat, full.dump_paths(b), "bounce %d paths (strips concatenated)" % b)


@pytest.mark.parametrize("bits,unfused", [(4, "0"), (2, "0"), (4, "1")], ids=["experimental-ray-sort-4bit", "experimental-ray-sort-2bit", "experimental-ray-sort-unfused"])
def test_ray_sort_changes_nothing(bits, unfused, monkeypatch):
    """PTD_PT_RAY_SORT (opt-in): rays are TRACED in (origin cell, direction octant) bin order, every record stays in its slot -
    PathSegments, ShadeableIntersections, live counts, the final partition layout and the G-buffer must all be bit-identical to
    the default scheduling, on a mesh scene (BVH) with the geoms in play too."""
    capi = _capi()
    monkeypatch.setenv("PTD_PT_RAY_SORT_BITS", str(bits))
    monkeypatch.setenv("PTD_PT_RAY_SORT_FROM", "1")                    # bin every bounce after the camera rays (default: from bounce 2)
    monkeypatch.setenv("PTD_PT_RAY_SORT_UNFUSED", unfused)             # 0: pt_shade writes the keys / histogram; 1: separate ray_bin_hist pass
    monkeypatch.delenv("PTD_PT_RAY_SORT", raising=False)
    sc = capi.Scene(path=os.path.join(SCENES, "hall_64x48.txt"))
    sc.set_resolution(160, 96)
    out = []
    for flags in (0, capi.PT_RAY_SORT):
        pt = capi.PathTracer(sc, flags=flags | capi.PT_TRACE | capi.PT_KEEP_TERMINATED)
        for rep in range(2):                                       # the second frame reuses the (re-zeroed) bin histograms
            g = pt.render_host()
        counts, run = pt.live_counts()
        out.append((g, counts[:run], [(pt.dump_paths(b), pt.dump_intersections(b)) for b in range(run)], pt.dump_final_paths(), pt.launches()))
    a, b = out
    assert a[0].tobytes() == b[0].tobytes() and a[1] == b[1]
    for k, ((pa, ia), (pb, ib)) in enumerate(zip(a[2], b[2])):
        _same(pa, pb, "bounce %d paths" % k)
        _same(ia, ib, "bounce %d intersections" % k)
    _same(a[3], b[3], "final partition layout")
    assert b[4] == a[4] + (3 if unfused == "1" else 2) * (sc.counts()[3] - 1)   # scan + scatter (+ the key pass when not fused) per bounce >= 1


@pytest.mark.parametrize("res", [(160, 96), (800, 600), (1000, 803)], ids=["30tiles", "938tiles", "1569tiles-ragged"])
def test_shade_kernel_variants_change_nothing(res, monkeypatch):
    """pt_shade comes as the pipelined persistent kernel (default: bulk-copy prefetch of the next tile, look-back on its own warp, stores one
    step later), as one tile per block (PTD_PT_SHADE_TILED=1) and that with a block-wide look-back (PTD_PT_WIDE_LOOKBACK=1).  The compaction
    must stay the same stable compaction: identical live counts, PathSegment order and G-buffer - with fewer tiles than there are
    persistent blocks, with several tiles per block, and with a ragged last tile."""
    capi = _capi()
    sc = capi.Scene(path=os.path.join(SCENES, "hall_64x48.txt"))
    sc.set_resolution(*res)
    out = []
    for env in ({}, {"PTD_PT_SHADE_TILED": "1"}, {"PTD_PT_WIDE_LOOKBACK": "1"}, {"PTD_SHADE_BLOCKS_PER_SM": "1"}):
        for k in ("PTD_PT_SHADE_TILED", "PTD_PT_WIDE_LOOKBACK", "PTD_SHADE_BLOCKS_PER_SM"):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        pt = capi.PathTracer(sc, flags=capi.PT_TRACE)
        for rep in range(2):
            g = pt.render_host()
        counts, run = pt.live_counts()
        out.append((g, counts[:run], [pt.dump_paths(b) for b in range(run)]))
    a = out[0]
    for b in out[1:]:
        assert a[1] == b[1] and a[0].tobytes() == b[0].tobytes()
        for k, (pa, pb) in enumerate(zip(a[2], b[2])):
            _same(pa, pb, "bounce %d paths" % k)


@pytest.mark.parametrize("combo", [{"PTD_PT_SMEM_STACK": "1"}, {"PTD_PT_SMEM_STACK": "1", "PTD_PT_RAY_SORT": "1", "PTD_PT_RAY_SORT_FROM": "1"}],
                         ids=["experimental-smem-stack", "experimental-smem-stack-with-ray-sort"])
def test_smem_stack_changes_nothing(combo, monkeypatch):
    """PTD_PT_SMEM_STACK=1 (opt-in): the top 16 entries of every lane's BVH traversal stack live in shared memory.  Same traversal,
    same records - alone and together with the ray binning."""
    capi = _capi()
    sc = capi.Scene(path=os.path.join(SCENES, "hall_64x48.txt"))
    sc.set_resolution(320, 200)
    for k in ("PTD_PT_SMEM_STACK", "PTD_PT_RAY_SORT", "PTD_PT_RAY_SORT_FROM"):
        monkeypatch.delenv(k, raising=False)
    ref = capi.PathTracer(sc, flags=capi.PT_TRACE)
    g_ref = ref.render_host()
    for k, v in combo.items():
        monkeypatch.setenv(k, v)
    pt = capi.PathTracer(sc, flags=capi.PT_TRACE)
    g = pt.render_host()
    assert g.tobytes() == g_ref.tobytes() and pt.live_counts() == ref.live_counts()
    for b in range(ref.live_counts()[1]):
        _same(pt.dump_paths(b), ref.dump_paths(b), "bounce %d paths" % b)
        _same(pt.dump_intersections(b), ref.dump_intersections(b), "bounce %d intersections" % b)


@pytest.mark.parametrize("mode", ["experimental-gated-mail"])
def test_row_strips_gated_mail(mode):
    """PTD_PT_GATED_MAIL (the flag the opt-in two-stream strip loop needs, PTD_STRIP_PIPELINE=1): the live-count mail is awaited by a
    one-warp gate kernel ahead of pt_shade.  Same bits as the untiled render, one extra launch per bounce >= 1 on the strips below
    the first."""
    capi = _capi()
    import torch
    W, H, n = 96, 80, 3
    sc = capi.Scene(path=os.path.join(SCENES, "hall_64x48.txt"))
    sc.set_resolution(W, H)
    ref = capi.PathTracer(sc).render_host()
    strips = [capi.PathTracer(sc, flags=capi.PT_GATED_MAIL, strip=r) for r in pt_strip_rows(capi, H, n)]
    infos = [s.export_info() for s in strips]
    for i, s in enumerate(strips):
        s.connect(infos, i)
    g = torch.zeros(10 * H * W, dtype=torch.float32, device="cuda")
    for rep in range(2):
        g.zero_()
        capi.PathTracer.render_group(strips, [g.data_ptr()] * n)
        torch.cuda.synchronize()
        assert g.cpu().numpy().reshape(10, H, W).tobytes() == ref.tobytes()


@pytest.mark.parametrize("repl", [None, "3"], ids=["tiled", "replicated"])
def test_strip_pipeline_equals_full_pipeline(tmp_path, repl, monkeypatch):
    """Path trace strips feeding denoiser strips (each strip only ever sees its own G-buffer rows) == the untiled frame loop."""
    capi = _capi()
    import torch
    from ai_path_tracer_denoiser_b200 import weights
    if repl:
        monkeypatch.setenv("PTD_DN_REPL_LEVEL", repl)          # levels >= 1/8 resolution replicated on every strip
    else:
        monkeypatch.setenv("PTD_DN_REPL_LEVEL", "6")                 # every level tiled
    wfile = weights.save_weights(weights.synthetic_state_dict(1234), str(tmp_path / "w.ptdw"))
    W, H, n = 96, 80, 3
    sc = capi.Scene(path=os.path.join(SCENES, "hall_64x48.txt"))
    sc.set_resolution(W, H)
    full_pt, full_dn = capi.PathTracer(sc), capi.Denoiser(wfile, H, W, flags=capi.DN_TF32)
    parts = [capi.strip_partition(H, n, i) for i in range(n)]
    pts = [capi.PathTracer(sc, strip=r) for r in pt_strip_rows(capi, H, n)]
    dns = [capi.Denoiser(wfile, H, W, flags=capi.DN_TF32, strip=p) for p in parts]
    pinfo, dinfo = [s.export_info() for s in pts], [s.export_info() for s in dns]
    for i in range(n):
        pts[i].connect(pinfo, i)
        dns[i].connect(dinfo, i)
    gs = [torch.zeros(10 * H * W, dtype=torch.float32, device="cuda") for _ in range(n)]     # one G-buffer per strip: only its rows are filled
    out = torch.zeros(3 * H * W, dtype=torch.float32, device="cuda")
    for k in range(3):
        cam = capi.frame_camera(sc.camera[0], k)
        ref = full_dn.forward_host(full_pt.render_host(cam), reset=(k == 0))
        capi.PathTracer.render_group(pts, [g.data_ptr() for g in gs], cam=cam)
        capi.Denoiser.forward_group(dns, [g.data_ptr() for g in gs], [out.data_ptr()] * n, k == 0)
        torch.cuda.synchronize()
        assert out.cpu().numpy().reshape(3, H, W).tobytes() == ref.tobytes(), k


def test_per_face_mtl_materials(tmp_path):
    """USEMTL extension: a mesh whose faces carry different materials (diffuse / mirror / glass / emitter from an
    MTL file).  BVH == brute force bit for bit, and the C oracle (which takes the per-face material ids as they are) agrees on the
    live counts and first-hit planes."""
    capi = _capi()
    from oracle import pt_oracle
    src = open(os.path.join(SCENES, "hall_small.obj")).read().splitlines()
    names = ["wall", "mirror", "glass", "lamp"]
    out, nf = ["mtllib hall.mtl"], 0
    for ln in src:
        if ln.startswith("f "):
            if nf % 97 == 0:
                out.append("usemtl %s" % names[(nf // 97) % 4 if (nf // 97) % 7 else 0])
            nf += 1
        out.append(ln)
    (tmp_path / "hall.obj").write_text("\n".join(out) + "\n")
    (tmp_path / "hall.mtl").write_text("newmtl wall\nKd .7 .7 .6\nillum 2\nnewmtl mirror\nKd .9 .9 .9\nKs .9 .9 .9\nillum 3\n"
                                       "newmtl glass\nKd 1 1 1\nKs 1 1 1\nNi 1.5\nillum 7\nnewmtl lamp\nKd 1 1 1\nKe 3 3 3\n")
    txt = open(os.path.join(SCENES, "hall_64x48.txt")).read()
    import re
    txt = re.sub(r"PATH \S+", "PATH %s" % (tmp_path / "hall.obj"), txt)
    txt = txt.rstrip() + "\nUSEMTL 1\n"                      # the MESH block is the file's last block
    (tmp_path / "hall_mtl.txt").write_text(txt)
    sc = capi.Scene(path=str(tmp_path / "hall_mtl.txt"))
    A = sc.arrays()
    assert len(np.unique(A["faces"]["mat"])) == 4
    cam = capi.frame_camera(sc.camera[0], 40)
    a = _render_ours(capi, A, cam, 0)
    b = _render_ours(capi, A, cam, capi.PT_NO_BVH)
    assert a["counts"] == b["counts"] and a["tensor"].tobytes() == b["tensor"].tobytes()
    for x, y in zip(a["trace"], b["trace"]):
        _same(x["paths"], y["paths"], "paths")
    ora = pt_oracle.render(A, cam, trace=True)
    for n_o, n_g in zip(a["counts"], ora["counts"]):
        assert abs(n_o - n_g) <= max(2, 0.005 * n_g)
    assert np.isclose(a["tensor"][3:], ora["tensor"][3:], rtol=0, atol=1e-4).mean() >= 0.999


def _ref_and_ours(capi, scene_name, cam_edit=None, depth=None, iter=1, res=None):
    """Render the same scene / camera with oracle A (reference kernels) and with libptd.so; returns (ref, ours, pt)."""
    from oracle import reflib
    if not reflib.available(""):
        pytest.fail("oracle/_ref/libref_pt.so missing on the GPU box")
    R = reflib.RefLib("")
    path = os.path.join(SCENES, scene_name)
    s = R.load_scene(path)
    sc = capi.Scene(path=path)
    cam = capi.frame_camera(sc.camera[0], 3)
    if res is not None:                                   # smaller than the file's 64x48, so the reference's buffers still fit
        sc.set_resolution(*res)
        cam = capi.frame_camera(sc.camera[0], 3)
    if cam_edit is not None:
        cam = cam_edit(cam)
    if depth is not None:
        R.set_depth(s, depth)
        sc.set_depth(depth)
    R.set_camera(s, cam)
    ref = R.gpu_render(s, iter=iter, trace=True)
    sc.set_camera(cam)
    pt = capi.PathTracer(sc, flags=capi.PT_TRACE | capi.PT_KEEP_TERMINATED)
    tensor = pt.render_host(iter=iter)
    counts, run = pt.live_counts()
    return ref, dict(tensor=tensor, counts=counts[:run], image=pt.dump_image(), final=pt.dump_final_paths(), pt=pt), pt


@pytest.mark.parametrize("scene", ["cornell_specular_64x48.txt", "hall_64x48.txt"])
def test_edge_odd_resolution_bit_exact(scene):
    """37 x 23: not a multiple of any tile size (ragged last tile in every kernel)."""
    capi = _capi()
    ref, ours, pt = _ref_and_ours(capi, scene, res=(37, 23))
    assert [b["n"] for b in ref["trace"]] == ours["counts"]
    for b in range(len(ours["counts"])):
        _same(pt.dump_paths(b), ref["trace"][b]["paths"], "bounce %d paths" % b)
    assert ours["tensor"].tobytes() == ref["tensor"].tobytes() and ours["image"].tobytes() == ref["image"].tobytes()
    _same(ours["final"], ref["final_paths"], "final paths")


def test_edge_depth_one_and_all_rays_miss():
    capi = _capi()
    ref, ours, pt = _ref_and_ours(capi, "cornell_64x48.txt", depth=1)
    assert ours["counts"] == [64 * 48] and ours["tensor"].tobytes() == ref["tensor"].tobytes()

    def look_away(cam):                                   # turn the camera around: nothing but the void in view
        cam = cam.copy()
        cam["view"] = -cam["view"]
        cam["right"] = -cam["right"]
        return cam
    ref, ours, pt = _ref_and_ours(capi, "cornell_64x48.txt", cam_edit=look_away)
    assert [b["n"] for b in ref["trace"]] == ours["counts"] == [64 * 48]      # every path dies at bounce 0: the next bounce sees n == 0
    assert ours["tensor"].tobytes() == ref["tensor"].tobytes() and not ours["tensor"].any()
    _same(ours["final"], ref["final_paths"], "final paths")


def test_edge_second_iteration_and_accumulation():
    """iter == 2: another RNG stream, no first-hit planes (pathtrace.cu:295,379 guards), radiance = image / 2; and iter 1 followed
    by iter 2 on one handle accumulates like the reference's dev_image."""
    capi = _capi()
    ref2, ours2, _ = _ref_and_ours(capi, "cornell_specular_64x48.txt", iter=2)
    assert ours2["tensor"].tobytes() == ref2["tensor"].tobytes() and ours2["image"].tobytes() == ref2["image"].tobytes()
    assert not ours2["tensor"][3:].any()
    ref1, ours1, pt = _ref_and_ours(capi, "cornell_specular_64x48.txt", iter=1)
    t12 = pt.render_host(iter=2)                          # same handle: accumulates onto iteration 1's image
    acc = (ref1["image"] + ref2["image"]).astype(np.float32)
    assert pt.dump_image().tobytes() == acc.tobytes()
    expect = (acc / np.float32(2)).astype(np.float32).reshape(48, 64, 3)[:, ::-1].transpose(2, 0, 1)     # x-mirrored planes 0-2
    assert t12[:3].tobytes() == np.ascontiguousarray(expect).tobytes()
    assert t12[3:].tobytes() == ours1["tensor"][3:].tobytes()       # first-hit planes keep iteration 1's values
