"""TEST INFRASTRUCTURE ONLY - ctypes binding of oracle/_ref/libref_pt*.so (see oracle/ref_driver.cu).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  The product package never does.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

# Record layouts of Inference/src/sceneStructs.h (sizes verified by ref_sizeof()).
PATH_DT = np.dtype([("o", "<f4", 3), ("d", "<f4", 3), ("color", "<f4", 3), ("pix", "<i4"), ("rb", "<i4")])          # :77-82, 44 B
ISX_DT = np.dtype([("t", "<f4"), ("n", "<f4", 3), ("mat", "<i4"), ("inside", "u1"), ("pad", "u1", 3), ("p", "<f4", 3)])  # :91-97, 36 B
GEOM_DT = np.dtype([("type", "<i4"), ("mat", "<i4"), ("trans", "<f4", 3), ("rot", "<f4", 3), ("scale", "<f4", 3),
                    ("T", "<f4", 16), ("invT", "<f4", 16), ("invTr", "<f4", 16), ("vel", "<f4", 3)])                  # :20-30, 248 B
FACE_DT = np.dtype([("v", "<f4", (3, 3)), ("n", "<f4", (3, 3)), ("mat", "<i4")])                                      # :40-44, 76 B
MAT_DT = np.dtype([("color", "<f4", 3), ("specex", "<f4"), ("speccolor", "<f4", 3), ("refl", "<f4"), ("refr", "<f4"),
                   ("ior", "<f4"), ("emit", "<f4")])                                                                 # :46-56, 44 B
CAM_DT = np.dtype([("res", "<i4", 2), ("pos", "<f4", 3), ("lookat", "<f4", 3), ("view", "<f4", 3), ("up", "<f4", 3),
                   ("right", "<f4", 3), ("fov", "<f4", 2), ("pixlen", "<f4", 2)])                                     # :58-67, 84 B
AABB_DT = np.dtype([("lb", "<f4", 3), ("ub", "<f4", 3)])                                                              # :84-87, 24 B
assert (PATH_DT.itemsize, ISX_DT.itemsize, GEOM_DT.itemsize, FACE_DT.itemsize, MAT_DT.itemsize, CAM_DT.itemsize,
        AABB_DT.itemsize) == (44, 36, 248, 76, 44, 84, 24)


def lib_path(variant=""):
    return os.path.join(HERE, "_ref", "libref_pt%s.so" % (("_" + variant) if variant else ""))


def available(variant=""):
    return os.path.exists(lib_path(variant))


class RefLib:
    """One loaded variant ("" = nvcc defaults, "nofma", "sort")."""

    def __init__(self, variant=""):
        self.lib = L = C.CDLL(lib_path(variant))
        L.ref_scene_load.restype = C.c_void_p
        L.ref_scene_load.argtypes = [C.c_char_p]
        for f in ("ref_scene_geoms", "ref_scene_materials", "ref_scene_faces", "ref_scene_meshbox", "ref_scene_camera"):
            getattr(L, f).restype = C.c_void_p
            getattr(L, f).argtypes = [C.c_void_p]
        L.ref_scene_counts.argtypes = [C.c_void_p, C.POINTER(C.c_int)]
        L.ref_scene_set_camera.argtypes = [C.c_void_p, C.c_void_p]
        L.ref_scene_set_depth.argtypes = [C.c_void_p, C.c_int]
        L.ref_trace_paths.restype = C.c_void_p
        L.ref_trace_isx.restype = C.c_void_p
        L.ref_gpu_render.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_float)]
        L.ref_cpu_render.restype = C.c_longlong
        L.ref_cpu_render.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_double)]
        L.ref_cpu_first_bounce_rows.restype = C.c_longlong
        L.ref_cpu_first_bounce_rows.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_double)]
        assert [L.ref_sizeof(i) for i in range(7)] == [44, 36, 248, 76, 44, 84, 24]

    # ---- scene -------------------------------------------------------------------------------
    def load_scene(self, path):
        cwd = os.getcwd()
        os.chdir(os.path.dirname(os.path.abspath(path)) or ".")   # MESH PATH entries are cwd-relative
        try:
            s = self.lib.ref_scene_load(os.path.abspath(path).encode())
        finally:
            os.chdir(cwd)
        if not s:
            raise RuntimeError("reference Scene() failed for %s" % path)
        return s

    def _arr(self, ptr, dt, n):
        if n == 0 or not ptr:
            return np.zeros(0, dt)
        buf = (C.c_char * (dt.itemsize * n)).from_address(ptr)
        return np.frombuffer(buf, dtype=dt, count=n).copy()

    def scene_arrays(self, s):
        cnt = (C.c_int * 5)()
        self.lib.ref_scene_counts(s, cnt)
        return dict(
            geoms=self._arr(self.lib.ref_scene_geoms(s), GEOM_DT, cnt[0]),
            materials=self._arr(self.lib.ref_scene_materials(s), MAT_DT, cnt[1]),
            faces=self._arr(self.lib.ref_scene_faces(s), FACE_DT, cnt[2]),
            mesh_box=self._arr(self.lib.ref_scene_meshbox(s), AABB_DT, 1),
            camera=self._arr(self.lib.ref_scene_camera(s), CAM_DT, 1),
            depth=cnt[3], iterations=cnt[4])

    def set_camera(self, s, cam):
        cam = np.ascontiguousarray(cam, dtype=CAM_DT)
        self.lib.ref_scene_set_camera(s, cam.ctypes.data)

    def set_depth(self, s, d):
        self.lib.ref_scene_set_depth(s, int(d))

    # ---- renders -----------------------------------------------------------------------------
    def _collect_trace(self):
        out = []
        for b in range(self.lib.ref_trace_bounces()):
            n = self.lib.ref_trace_count(b)
            out.append(dict(n=n, paths=self._arr(self.lib.ref_trace_paths(b), PATH_DT, n),
                            isx=self._arr(self.lib.ref_trace_isx(b), ISX_DT, n)))
        return out

    def _render(self, s, fn, trace, *extra):
        cam = self._arr(self.lib.ref_scene_camera(s), CAM_DT, 1)[0]
        W, H = int(cam["res"][0]), int(cam["res"][1])
        P = W * H
        tensor = np.zeros((10, H, W), np.float32)
        image = np.zeros((P, 3), np.float32)
        final = np.zeros(P, PATH_DT)
        self.lib.ref_trace_enable(1 if trace else 0)
        res = fn(tensor, image, final, *extra)
        tr = self._collect_trace() if trace else None
        self.lib.ref_trace_enable(0)
        self.lib.ref_trace_clear()
        return dict(tensor=tensor, image=image, final_paths=final, trace=tr, **res)

    def cpu_render(self, s, iter=1, sort_material=False, trace=False):
        """oracle B (reference host functions, restated loops)."""
        def fn(tensor, image, final):
            ms = C.c_double()
            total = self.lib.ref_cpu_render(s, iter, 1 if sort_material else 0, tensor.ctypes.data, image.ctypes.data,
                                            final.ctypes.data, C.byref(ms))
            return dict(sum_live=int(total), ms=ms.value)
        return self._render(s, fn, trace)

    def gpu_render(self, s, iter=1, trace=False):
        """oracle A (verbatim reference kernels on the GPU)."""
        def fn(tensor, image, final):
            ms = C.c_float()
            rc = self.lib.ref_gpu_render(s, iter, tensor.ctypes.data, image.ctypes.data, final.ctypes.data, C.byref(ms))
            if rc != 0:
                raise RuntimeError("reference GPU render failed, cuda error %d" % rc)
            return dict(ms=ms.value)
        return self._render(s, fn, trace)

    def cpu_first_bounce_rows(self, s, row_stride, iter=1):
        ms = C.c_double()
        rays = self.lib.ref_cpu_first_bounce_rows(s, iter, row_stride, C.byref(ms))
        return int(rays), ms.value
