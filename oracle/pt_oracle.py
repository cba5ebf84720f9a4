"""TEST INFRASTRUCTURE ONLY - ctypes binding of oracle/libpt_oracle.so (oracle/pt_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module; the product package never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from .reflib import PATH_DT, ISX_DT, GEOM_DT, FACE_DT, MAT_DT, CAM_DT, AABB_DT  # layouts of sceneStructs.h

HERE = os.path.dirname(os.path.abspath(__file__))
_TRACE_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p)
_lib = None


def lib():
    global _lib
    if _lib is None:
        so = os.path.join(HERE, "libpt_oracle.so")
        src = os.path.join(HERE, "pt_oracle.c")
        if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
            subprocess.check_call(["make", "-C", HERE, "libpt_oracle.so"], stdout=subprocess.DEVNULL)
        L = C.CDLL(so)
        L.pto_render.restype = C.c_longlong
        L.pto_render.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                 C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, _TRACE_FN, C.c_void_p]
        L.pto_hash_seed.restype = C.c_uint32
        L.pto_triangle.restype = C.c_float
        L.pto_triangle.argtypes = [C.c_void_p] * 5
        L.pto_camera_from_scene.argtypes = [C.c_void_p, C.c_float]
        L.pto_camera_orbit_params.argtypes = [C.c_void_p] * 4
        L.pto_camera_orbit.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_float]
        _lib = L
    return _lib


def camera_orbit_params(cam):
    cam = np.ascontiguousarray(cam, CAM_DT).reshape(1).copy()
    z, p, t = C.c_float(), C.c_float(), C.c_float()
    lib().pto_camera_orbit_params(cam.ctypes.data, C.byref(z), C.byref(p), C.byref(t))
    return z.value, p.value, t.value


def camera_orbit(cam, zoom, phi, theta):
    """main.cpp:126-138: returns a copy of `cam` with position/view/up/right rebuilt."""
    cam = np.ascontiguousarray(cam, CAM_DT).reshape(1).copy()
    lib().pto_camera_orbit(cam.ctypes.data, zoom, phi, theta)
    return cam


def frame_camera(cam, frame=0, dphi=0.002):
    """Camera of frame k of the pan: phi_k = phi_0 + dphi*k through runCuda's orbit maths."""
    zoom, phi, theta = camera_orbit_params(cam)
    return camera_orbit(cam, zoom, np.float32(phi + np.float32(dphi) * np.float32(frame)), theta)


def render(scene, cam, iter=1, sort_material=False, trace=False):
    """scene: dict(geoms, materials, faces, mesh_box, depth) of numpy record arrays; cam: CAM_DT record."""
    L = lib()
    cam = np.ascontiguousarray(cam, CAM_DT).reshape(1)
    W, H = int(cam["res"][0][0]), int(cam["res"][0][1])
    P = W * H
    geoms = np.ascontiguousarray(scene["geoms"], GEOM_DT)
    mats = np.ascontiguousarray(scene["materials"], MAT_DT)
    faces = np.ascontiguousarray(scene["faces"], FACE_DT)
    box = np.ascontiguousarray(scene["mesh_box"], AABB_DT).reshape(1)
    depth = int(scene["depth"])
    tensor = np.zeros((10, H, W), np.float32)
    image = np.zeros((P, 3), np.float32)
    final = np.zeros(P, PATH_DT)
    counts = np.zeros(max(depth, 1), np.int32)
    tr = []

    def cb(user, d, n, pp, ip):
        paths = np.frombuffer((C.c_char * (44 * n)).from_address(pp), PATH_DT, n).copy() if n else np.zeros(0, PATH_DT)
        isx = np.frombuffer((C.c_char * (36 * n)).from_address(ip), ISX_DT, n).copy() if n else np.zeros(0, ISX_DT)
        tr.append(dict(n=n, paths=paths, isx=isx))

    fn = _TRACE_FN(cb) if trace else C.cast(None, _TRACE_FN)
    total = L.pto_render(len(geoms), geoms.ctypes.data, len(mats), mats.ctypes.data, len(faces),
                         faces.ctypes.data if len(faces) else None, box.ctypes.data, cam.ctypes.data, depth, iter,
                         1 if sort_material else 0, tensor.ctypes.data, image.ctypes.data, final.ctypes.data,
                         counts.ctypes.data, fn, None)
    return dict(tensor=tensor, image=image, final_paths=final, counts=counts, sum_live=int(total), trace=tr if trace else None)
