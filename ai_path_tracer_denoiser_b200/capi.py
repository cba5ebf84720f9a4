"""ctypes binding of libptd.so - the C ABI declared in include/ptd.h.

This is the host-side mirror the tests and bench.py drive: same entry points a C++ caller (ptd_cli, or the
reference's runCuda()) binds.  There is no fallback of any kind: if the shared library
is missing the import of this module raises, and every compute call fails when no CUDA device exists.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PTD_LIBPTD") or os.path.join(HERE, "libptd.so")     # (PTD_LIBPTD: an alternative build, for A/B timing)

# record layouts of Inference/src/sceneStructs.h == include/ptd.h
PATH_DT = np.dtype([("o", "<f4", 3), ("d", "<f4", 3), ("color", "<f4", 3), ("pix", "<i4"), ("rb", "<i4")])
ISX_DT = np.dtype([("t", "<f4"), ("n", "<f4", 3), ("mat", "<i4"), ("inside", "u1"), ("pad", "u1", 3), ("p", "<f4", 3)])
GEOM_DT = np.dtype([("type", "<i4"), ("mat", "<i4"), ("trans", "<f4", 3), ("rot", "<f4", 3), ("scale", "<f4", 3),
                    ("T", "<f4", 16), ("invT", "<f4", 16), ("invTr", "<f4", 16), ("vel", "<f4", 3)])
FACE_DT = np.dtype([("v", "<f4", (3, 3)), ("n", "<f4", (3, 3)), ("mat", "<i4")])
MAT_DT = np.dtype([("color", "<f4", 3), ("specex", "<f4"), ("speccolor", "<f4", 3), ("refl", "<f4"), ("refr", "<f4"),
                   ("ior", "<f4"), ("emit", "<f4")])
CAM_DT = np.dtype([("res", "<i4", 2), ("pos", "<f4", 3), ("lookat", "<f4", 3), ("view", "<f4", 3), ("up", "<f4", 3),
                   ("right", "<f4", 3), ("fov", "<f4", 2), ("pixlen", "<f4", 2)])
AABB_DT = np.dtype([("lb", "<f4", 3), ("ub", "<f4", 3)])

PT_SORT_MATERIAL, PT_TRACE, PT_NO_BVH, PT_KEEP_TERMINATED, PT_GATED_MAIL, PT_RAY_SORT = 1, 2, 4, 8, 16, 32
DN_FP32, DN_TF32, DN_3XTF32, DN_F16, DN_FP32_BATCH_STATS, DN_2XF16 = 0, 1, 2, 3, 4, 5

EXPORTS = """ptd_last_error ptd_version ptd_sizeof ptd_device_count ptd_scene_load ptd_scene_from_arrays ptd_scene_free
ptd_scene_counts ptd_scene_geoms ptd_scene_materials ptd_scene_faces ptd_scene_mesh_box ptd_scene_camera
ptd_scene_set_resolution ptd_scene_set_depth ptd_camera_orbit_params ptd_camera_orbit ptd_pt_create ptd_pt_destroy
ptd_pt_render ptd_pt_render_host ptd_pt_export_rgba8 ptd_pt_live_counts ptd_pt_dump_paths ptd_pt_dump_intersections
ptd_pt_dump_final_paths ptd_pt_dump_image ptd_pt_bvh_stats ptd_dn_create ptd_dn_destroy ptd_dn_forward
ptd_dn_forward_host ptd_dn_create_strip ptd_dn_padded_size ptd_dn_dump_hidden ptd_dn_launches_per_forward
ptd_pt_launches_last_render ptd_dn_profile ptd_dn_launch_times ptd_dn_launch_name ptd_pt_profile
ptd_pt_launch_times ptd_dn_strip_partition ptd_dn_strip_info_size ptd_dn_strip_export ptd_dn_strip_connect
ptd_dn_forward_group ptd_pt_create_strip ptd_pt_strip_info_size ptd_pt_strip_export ptd_pt_strip_connect
ptd_pt_render_group ptd_frame_host ptd_frame_submit ptd_frame_wait ptd_frame_slots ptd_frame_timer ptd_bvh_probe ptd_bvh_probe_order""".split()


class PtdError(RuntimeError):
    pass


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            # a fresh checkout (built artefacts are git-ignored): build in-tree once - nvcc cross-compiles sm_100a without a GPU
            import subprocess
            try:
                subprocess.check_call(["make", "-C", os.path.join(HERE, "csrc"), "-j4", "all"], stdout=subprocess.DEVNULL)
            except Exception as e:
                raise PtdError("libptd.so is not built (%s) and `make -C ai_path_tracer_denoiser_b200/csrc` failed: %s - "
                               "there is no CPU fallback" % (LIB_PATH, e))
        L = C.CDLL(LIB_PATH)
        L.ptd_last_error.restype = C.c_char_p
        for f in ("ptd_scene_geoms", "ptd_scene_materials", "ptd_scene_faces", "ptd_scene_mesh_box", "ptd_scene_camera"):
            getattr(L, f).restype = C.c_void_p
            getattr(L, f).argtypes = [C.c_void_p]
        L.ptd_scene_load.argtypes = [C.c_char_p, C.POINTER(C.c_void_p)]
        L.ptd_scene_from_arrays.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                            C.c_int, C.c_int, C.POINTER(C.c_void_p)]
        L.ptd_scene_free.argtypes = [C.c_void_p]
        L.ptd_scene_free.restype = None
        L.ptd_scene_counts.argtypes = [C.c_void_p, C.POINTER(C.c_int)]
        L.ptd_scene_set_resolution.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.ptd_scene_set_depth.argtypes = [C.c_void_p, C.c_int]
        L.ptd_camera_orbit_params.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float)]
        L.ptd_camera_orbit.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_float]
        L.ptd_pt_create.argtypes = [C.c_void_p, C.c_int, C.c_uint, C.POINTER(C.c_void_p)]
        L.ptd_pt_destroy.argtypes = [C.c_void_p]
        L.ptd_pt_destroy.restype = None
        L.ptd_pt_render.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.ptd_pt_render_host.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.ptd_frame_submit.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.ptd_frame_wait.argtypes = [C.c_void_p]
        L.ptd_frame_slots.argtypes = []
        L.ptd_frame_slots.restype = C.c_int
        L.ptd_frame_timer.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_float)]
        L.ptd_frame_host.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.ptd_bvh_probe_order.argtypes = [C.c_void_p, C.c_int, C.c_uint, C.c_int, C.POINTER(C.c_double)]
        L.ptd_bvh_probe.argtypes = [C.c_void_p, C.c_int, C.c_uint, C.c_int, C.POINTER(C.c_double)]
        L.ptd_pt_export_rgba8.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.ptd_pt_live_counts.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.c_int, C.POINTER(C.c_int)]
        L.ptd_pt_dump_paths.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.POINTER(C.c_int)]
        L.ptd_pt_dump_intersections.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.POINTER(C.c_int)]
        L.ptd_pt_dump_final_paths.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.ptd_pt_dump_image.argtypes = [C.c_void_p, C.c_void_p]
        L.ptd_pt_bvh_stats.argtypes = [C.c_void_p] + [C.POINTER(C.c_int)] * 4
        L.ptd_pt_launches_last_render.argtypes = [C.c_void_p]
        L.ptd_dn_create.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_uint, C.POINTER(C.c_void_p)]
        L.ptd_dn_destroy.argtypes = [C.c_void_p]
        L.ptd_dn_destroy.restype = None
        L.ptd_dn_forward.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.ptd_dn_forward_host.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.ptd_dn_padded_size.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.ptd_dn_dump_hidden.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.ptd_dn_launches_per_forward.argtypes = [C.c_void_p]
        L.ptd_dn_profile.argtypes = [C.c_void_p, C.c_int]
        L.ptd_dn_launch_times.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_int)]
        L.ptd_dn_launch_name.argtypes = [C.c_void_p, C.c_int]
        L.ptd_dn_launch_name.restype = C.c_char_p
        L.ptd_pt_profile.argtypes = [C.c_void_p, C.c_int]
        L.ptd_pt_launch_times.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_int)]
        L.ptd_pt_create_strip.argtypes = [C.c_void_p, C.c_int, C.c_uint, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
        L.ptd_pt_strip_export.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.ptd_pt_strip_connect.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        L.ptd_pt_render_group.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.ptd_dn_create_strip.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint, C.POINTER(C.c_void_p)]
        L.ptd_dn_strip_partition.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.ptd_dn_strip_export.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.ptd_dn_strip_connect.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        L.ptd_dn_forward_group.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        _lib = L
    return _lib


def check(rc, what=""):
    if rc != 0:
        raise PtdError("%s failed (%d): %s" % (what or "ptd call", rc, lib().ptd_last_error().decode()))


def device_count():
    return lib().ptd_device_count()


def frame_slots():
    """Frames ptd_frame_submit keeps in flight (3): submit frame k, then wait for frame k - (frame_slots() - 1)."""
    return lib().ptd_frame_slots()


def _view(ptr, dt, n):
    if n == 0 or not ptr:
        return np.zeros(0, dt)
    buf = (C.c_char * (dt.itemsize * n)).from_address(ptr)
    return np.frombuffer(buf, dtype=dt, count=n).copy()


class Scene:
    """`new Scene(file)` of the reference (scene.cpp:11-42) / or built from record arrays."""

    def __init__(self, path=None, arrays=None):
        L = lib()
        self.h = C.c_void_p()
        if path is not None:
            check(L.ptd_scene_load(os.fsencode(path), C.byref(self.h)), "ptd_scene_load(%s)" % path)
        else:
            a = arrays
            geoms = np.ascontiguousarray(a["geoms"], GEOM_DT)
            mats = np.ascontiguousarray(a["materials"], MAT_DT)
            faces = np.ascontiguousarray(a["faces"], FACE_DT)
            box = np.ascontiguousarray(a["mesh_box"], AABB_DT).reshape(-1)
            cam = np.ascontiguousarray(a["camera"], CAM_DT).reshape(-1)
            check(L.ptd_scene_from_arrays(len(geoms), geoms.ctypes.data if len(geoms) else None, len(mats),
                                          mats.ctypes.data if len(mats) else None, len(faces),
                                          faces.ctypes.data if len(faces) else None,
                                          box.ctypes.data if len(box) else None, cam.ctypes.data, int(a["depth"]),
                                          int(a.get("iterations", 1)), C.byref(self.h)), "ptd_scene_from_arrays")

    def __del__(self):
        try:
            if getattr(self, "h", None):
                lib().ptd_scene_free(self.h)
                self.h = None
        except Exception:      # interpreter shutdown: module globals may already be gone
            pass

    def counts(self):
        c = (C.c_int * 5)()
        check(lib().ptd_scene_counts(self.h, c))
        return list(c)

    def arrays(self):
        L = lib()
        c = self.counts()
        return dict(geoms=_view(L.ptd_scene_geoms(self.h), GEOM_DT, c[0]), materials=_view(L.ptd_scene_materials(self.h), MAT_DT, c[1]),
                    faces=_view(L.ptd_scene_faces(self.h), FACE_DT, c[2]), mesh_box=_view(L.ptd_scene_mesh_box(self.h), AABB_DT, 1),
                    camera=_view(L.ptd_scene_camera(self.h), CAM_DT, 1), depth=c[3], iterations=c[4])

    @property
    def camera(self):
        return _view(lib().ptd_scene_camera(self.h), CAM_DT, 1)

    def set_camera(self, cam):
        cam = np.ascontiguousarray(cam, CAM_DT).reshape(1)
        C.memmove(lib().ptd_scene_camera(self.h), cam.ctypes.data, CAM_DT.itemsize)

    def set_resolution(self, W, H):
        check(lib().ptd_scene_set_resolution(self.h, W, H))

    def set_depth(self, d):
        check(lib().ptd_scene_set_depth(self.h, d))


def camera_orbit_params(cam):
    cam = np.ascontiguousarray(cam, CAM_DT).reshape(1).copy()
    z, p, t = C.c_float(), C.c_float(), C.c_float()
    check(lib().ptd_camera_orbit_params(cam.ctypes.data, C.byref(z), C.byref(p), C.byref(t)))
    return z.value, p.value, t.value


def camera_orbit(cam, zoom, phi, theta):
    cam = np.ascontiguousarray(cam, CAM_DT).reshape(1).copy()
    check(lib().ptd_camera_orbit(cam.ctypes.data, zoom, phi, theta))
    return cam


def frame_camera(cam, frame=0, dphi=0.002):
    """Camera of frame k of the pan: phi_k = phi_0 + dphi*k through runCuda's orbit maths."""
    zoom, phi, theta = camera_orbit_params(cam)
    return camera_orbit(cam, zoom, float(np.float32(phi + np.float32(dphi) * np.float32(frame))), theta)


class PathTracer:
    """pathtraceInit / pathtrace / pathtraceFree (pathtrace.h:6-8) behind ptd_pt_*."""

    def __init__(self, scene, device=0, flags=0, strip=None):
        """strip=(row0, rows): trace only those image rows (ptd_pt_create_strip, multi-GPU tiling)."""
        self.scene = scene
        self.h = C.c_void_p()
        cam = scene.camera[0]
        self.W, self.H = int(cam["res"][0]), int(cam["res"][1])
        if strip is None:
            check(lib().ptd_pt_create(scene.h, device, flags, C.byref(self.h)), "ptd_pt_create")
            self.P = self.W * self.H
        else:
            check(lib().ptd_pt_create_strip(scene.h, device, flags, strip[0], strip[1], C.byref(self.h)), "ptd_pt_create_strip")
            self.P = self.W * strip[1]
        self.strip = strip
        self.depth = scene.counts()[3]

    def export_info(self):
        n = lib().ptd_pt_strip_info_size()
        buf = C.create_string_buffer(n)
        check(lib().ptd_pt_strip_export(self.h, buf, n), "ptd_pt_strip_export")
        return buf.raw

    def connect(self, infos, my_rank):
        """infos: the exported blobs of ALL strips in strip order (top first)."""
        check(lib().ptd_pt_strip_connect(self.h, b"".join(infos), len(infos), my_rank), "ptd_pt_strip_connect")

    @staticmethod
    def render_group(strips, gbuf_ptrs, cam=None, iter=1, streams=None):
        """Same-process strip group: one iteration, issued bounce by bounce over the handles (ptd_pt_render_group)."""
        n = len(strips)
        hs = (C.c_void_p * n)(*[s.h for s in strips])
        g = (C.c_void_p * n)(*gbuf_ptrs)
        st = (C.c_void_p * n)(*streams) if streams else None
        camp = None
        if cam is not None:
            cam = np.ascontiguousarray(cam, CAM_DT).reshape(1)
            camp = cam.ctypes.data
        check(lib().ptd_pt_render_group(hs, n, camp, iter, g, st), "ptd_pt_render_group")

    def __del__(self):
        try:
            if getattr(self, "h", None):
                lib().ptd_pt_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def render_host(self, cam=None, iter=1):
        """== pathtrace(pbo, 0, iter) + the D2H of host_tensor (pathtrace.cu:525). Returns float32 [10,H,W]."""
        out = np.zeros((10, self.H, self.W), np.float32)
        camp = None
        if cam is not None:
            cam = np.ascontiguousarray(cam, CAM_DT).reshape(1)
            camp = cam.ctypes.data
        check(lib().ptd_pt_render_host(self.h, camp, iter, out.ctypes.data), "ptd_pt_render_host")
        return out

    def frame_host(self, dn, cam=None, iter=1, reset=False, want_gbuffer=True):
        """One frame of runCuda()'s body through ptd_frame_host: (G-buffer [10,H,W] or None, denoised frame [3,H,W])."""
        g = np.zeros((10, self.H, self.W), np.float32) if want_gbuffer else None
        rgb = np.zeros((3, self.H, self.W), np.float32)
        camp = None
        if cam is not None:
            cam = np.ascontiguousarray(cam, CAM_DT).reshape(1)
            camp = cam.ctypes.data
        check(lib().ptd_frame_host(self.h, dn.h, camp, iter, 1 if reset else 0, g.ctypes.data if want_gbuffer else None, rgb.ctypes.data), "ptd_frame_host")
        return g, rgb

    def frame_submit(self, dn, rgb_out, gbuf_out=None, cam=None, iter=1, reset=False):
        """ptd_frame_submit: enqueue one frame; rgb_out [3,H,W] / gbuf_out [10,H,W] are float32 numpy arrays (or anything with .ctypes /
        .data_ptr()) that must stay alive until the matching frame_wait()."""
        def ptr(a):
            return None if a is None else (C.c_void_p(a.data_ptr()) if hasattr(a, "data_ptr") else a.ctypes.data)
        camp = None
        if cam is not None:
            cam = np.ascontiguousarray(cam, CAM_DT).reshape(1)
            camp = cam.ctypes.data
        check(lib().ptd_frame_submit(self.h, dn.h, camp, iter, 1 if reset else 0, ptr(gbuf_out), ptr(rgb_out)), "ptd_frame_submit")

    def frame_wait(self):
        check(lib().ptd_frame_wait(self.h), "ptd_frame_wait")

    def frame_timer_start(self):
        check(lib().ptd_frame_timer(self.h, 0, None), "ptd_frame_timer")

    def frame_timer_stop(self):
        ms = C.c_float()
        check(lib().ptd_frame_timer(self.h, 1, C.byref(ms)), "ptd_frame_timer")
        return ms.value

    def render(self, gbuf_dev_ptr, cam=None, iter=1, stream=None):
        camp = None
        if cam is not None:
            cam = np.ascontiguousarray(cam, CAM_DT).reshape(1)
            camp = cam.ctypes.data
        check(lib().ptd_pt_render(self.h, camp, iter, gbuf_dev_ptr, stream), "ptd_pt_render")

    def live_counts(self):
        c = (C.c_int * self.depth)()
        run = C.c_int()
        check(lib().ptd_pt_live_counts(self.h, c, self.depth, C.byref(run)))
        return list(c), run.value

    def dump_paths(self, bounce):
        buf = np.zeros(self.P, PATH_DT)
        n = C.c_int()
        check(lib().ptd_pt_dump_paths(self.h, bounce, buf.ctypes.data, self.P, C.byref(n)))
        return buf[:n.value].copy()

    def dump_intersections(self, bounce):
        buf = np.zeros(self.P, ISX_DT)
        n = C.c_int()
        check(lib().ptd_pt_dump_intersections(self.h, bounce, buf.ctypes.data, self.P, C.byref(n)))
        return buf[:n.value].copy()

    def dump_final_paths(self):
        buf = np.zeros(self.P, PATH_DT)
        check(lib().ptd_pt_dump_final_paths(self.h, buf.ctypes.data, self.P))
        return buf

    def dump_image(self):
        buf = np.zeros((self.P, 3), np.float32)
        check(lib().ptd_pt_dump_image(self.h, buf.ctypes.data))
        return buf

    def bvh_stats(self):
        v = [C.c_int() for _ in range(4)]
        check(lib().ptd_pt_bvh_stats(self.h, *[C.byref(x) for x in v]))
        return dict(nodes=v[0].value, leaves=v[1].value, max_leaf=v[2].value, max_depth=v[3].value)

    def launches(self):
        return lib().ptd_pt_launches_last_render(self.h)

    def profile(self, on=True):
        check(lib().ptd_pt_profile(self.h, 1 if on else 0))

    def launch_times(self):
        ms = np.zeros(4096, np.float32)
        n = C.c_int()
        check(lib().ptd_pt_launch_times(self.h, ms.ctypes.data, ms.size, C.byref(n)))
        return ms[:n.value].copy()


def strip_partition(H, nstrips, index):
    """(row0, rows) of strip `index`: the padded frame's 32-row groups split as evenly as possible."""
    a, b = C.c_int(), C.c_int()
    check(lib().ptd_dn_strip_partition(H, nstrips, index, C.byref(a), C.byref(b)), "ptd_dn_strip_partition")
    return a.value, b.value


class Denoiser:
    """network_prediction_faster_version (main.cpp:101-118) / AutoEncoder.forward(x, j) behind ptd_dn_*.
    strip=(row0, rows): a row-strip handle of the multi-GPU tiling (ptd_dn_create_strip)."""

    def __init__(self, weights_path, H, W, device=0, flags=DN_TF32, strip=None):
        self.h = C.c_void_p()
        if strip is None:
            check(lib().ptd_dn_create(os.fsencode(weights_path), H, W, device, flags, C.byref(self.h)), "ptd_dn_create")
        else:
            check(lib().ptd_dn_create_strip(os.fsencode(weights_path), H, W, strip[0], strip[1], device, flags, C.byref(self.h)), "ptd_dn_create_strip")
        self.H, self.W, self.strip = H, W, strip

    def export_info(self):
        """POD blob describing this strip's arena (bytes); all-gather it and pass the neighbours' blobs to connect()."""
        n = lib().ptd_dn_strip_info_size()
        buf = C.create_string_buffer(n)
        check(lib().ptd_dn_strip_export(self.h, buf, n), "ptd_dn_strip_export")
        return buf.raw

    def connect(self, infos, my_rank):
        """infos: the exported blobs of ALL strips in strip order (top first)."""
        check(lib().ptd_dn_strip_connect(self.h, b"".join(infos), len(infos), my_rank), "ptd_dn_strip_connect")

    def __del__(self):
        try:
            if getattr(self, "h", None):
                lib().ptd_dn_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def forward_host(self, gbuf, reset):
        gbuf = np.ascontiguousarray(gbuf, np.float32)
        assert gbuf.shape == (10, self.H, self.W)
        out = np.empty((3, self.H, self.W), np.float32)
        check(lib().ptd_dn_forward_host(self.h, gbuf.ctypes.data, out.ctypes.data, 1 if reset else 0), "ptd_dn_forward_host")
        return out

    def forward(self, gbuf_dev_ptr, rgb_dev_ptr, reset, stream=None):
        check(lib().ptd_dn_forward(self.h, gbuf_dev_ptr, rgb_dev_ptr, 1 if reset else 0, stream), "ptd_dn_forward")

    @staticmethod
    def forward_group(strips, gbuf_ptrs, rgb_ptrs, reset, streams=None):
        """Same-process strip group: one frame, issued layer by layer over the handles (ptd_dn_forward_group)."""
        n = len(strips)
        hs = (C.c_void_p * n)(*[s.h for s in strips])
        g = (C.c_void_p * n)(*gbuf_ptrs)
        r = (C.c_void_p * n)(*rgb_ptrs)
        st = (C.c_void_p * n)(*streams) if streams else None
        check(lib().ptd_dn_forward_group(hs, n, g, r, 1 if reset else 0, st), "ptd_dn_forward_group")

    def padded_size(self):
        a, b = C.c_int(), C.c_int()
        check(lib().ptd_dn_padded_size(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def dump_hidden(self, level):
        Hp, Wp = self.padded_size()
        cap = 101 * Hp * Wp
        c, h, w = C.c_int(), C.c_int(), C.c_int()
        buf = np.zeros(cap >> (2 * level) if level else cap, np.float32)
        check(lib().ptd_dn_dump_hidden(self.h, level, buf.ctypes.data, buf.size, C.byref(c), C.byref(h), C.byref(w)))
        return buf[:c.value * h.value * w.value].reshape(c.value, h.value, w.value).copy()

    def launches(self):
        return lib().ptd_dn_launches_per_forward(self.h)

    def profile(self, on=True):
        check(lib().ptd_dn_profile(self.h, 1 if on else 0))

    def launch_times(self):
        """[(name, ms)] of the last profiled forward, in launch order."""
        ms = np.zeros(256, np.float32)
        n = C.c_int()
        check(lib().ptd_dn_launch_times(self.h, ms.ctypes.data, ms.size, C.byref(n)))
        return [(lib().ptd_dn_launch_name(self.h, i).decode(), float(ms[i])) for i in range(n.value)]
