"""ctypes access to the reference's own GPU library built by oracle/build_ref.py
(oracle/_ref/libgpu_ref*.so = /root/reference/.../gpu_process.cu compiled unmodified against the
stand-in Eigen header + oracle/ref_harness.cu).  TEST INFRASTRUCTURE."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "oracle", "_ref")


def available(nofma: bool = True) -> bool:
    return os.path.exists(os.path.join(REF_DIR, "libgpu_ref_nofma.so" if nofma else "libgpu_ref.so"))


def _p(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


class RefMap:
    """The reference keeps ONE map per process in __device__ globals (gpu_process.cu:20-56):
    create one RefMap per loaded library at a time."""
    _libs = {}

    def __init__(self, length, resolution, mahalanobis=2.5, obstacle_threshold=0.7, nofma=True):
        name = "libgpu_ref_nofma.so" if nofma else "libgpu_ref.so"
        if name not in RefMap._libs:
            RefMap._libs[name] = C.CDLL(os.path.join(REF_DIR, name))
        self.lib = RefMap._libs[name]
        self.L, self.res = int(length), float(resolution)
        self.lib.ref_init(self.L, C.c_float(self.res), C.c_float(mahalanobis), C.c_float(obstacle_threshold))

    def move(self, pos):
        p = np.asarray(pos, np.float32)
        centre, start, shift = np.zeros(2, np.float32), np.zeros(2, np.int32), np.zeros(2, np.float32)
        self.lib.ref_move(_p(p), _p(centre), _p(start), _p(shift), C.c_float(self.res))
        return centre, start, shift

    def process_points(self, x, y, z, frame):
        x, y, z = (np.array(a, np.float32, copy=True) for a in (x, y, z))
        n = x.shape[0]
        key = np.empty(n, np.int32)
        var, xt, yt, zt = (np.empty(n, np.float32) for _ in range(4))
        arr = lambda v: np.array(v[:], np.float32)
        T, sJ, rv = arr(frame.T), arr(frame.sensor_jacobian), arr(frame.rotation_variance)
        cs, pm, bs = arr(frame.C_SB_transpose), arr(frame.P_mul_C_BM_transpose), arr(frame.B_r_BS_skew)
        s = frame.sensor
        self.lib.ref_process_points(_p(key), _p(x), _p(y), _p(z), _p(var), _p(xt), _p(yt), _p(zt), _p(T), n,
                                    C.c_double(frame.rel_lower), C.c_double(frame.rel_upper), C.c_float(s.min_radius),
                                    C.c_float(s.beam_angle), C.c_float(s.beam_constant), _p(sJ), _p(rv), _p(cs), _p(pm),
                                    _p(bs))
        return key, var, xt, yt, zt

    def fuse_points(self, index, R, G, B, intensity, height, var):
        i32 = lambda a: np.ascontiguousarray(a, np.int32)
        f32 = lambda a: np.ascontiguousarray(a, np.float32)
        index, R, G, B = i32(index), i32(R), i32(G), i32(B)
        intensity, height, var = f32(intensity), f32(height), f32(var)
        self.lib.ref_fuse(index.shape[0], _p(index), _p(R), _p(G), _p(B), _p(intensity), _p(height), _p(var))

    def var_update(self, dv):
        self.lib.ref_var_update(C.c_float(dv))

    def map_feature(self):
        n = self.L * self.L
        out = {k: np.zeros(n, np.float32) for k in ("elevation", "variance", "rough", "slope", "traver", "intensity")}
        out.update({k: np.zeros(n, np.int32) for k in ("color_r", "color_g", "color_b")})
        self.lib.ref_map_feature(_p(out["elevation"]), _p(out["variance"]), _p(out["color_r"]), _p(out["color_g"]),
                                 _p(out["color_b"]), _p(out["rough"]), _p(out["slope"]), _p(out["traver"]),
                                 _p(out["intensity"]))
        return out

    def raytracing(self):
        self.lib.ref_raytracing()

    def get_layer(self, name):
        which = {"lowest": 0, "traver": 1, "elevation": 2, "variance": 3}[name]
        out = np.empty(self.L * self.L, np.float32)
        self.lib.ref_get_layer(which, _p(out))
        return out.reshape(self.L, self.L)

    def set_layer(self, name, arr):
        which = {"lowest": 0, "traver": 1, "elevation": 2, "variance": 3}[name]
        a = np.ascontiguousarray(arr, np.float32).reshape(-1)
        self.lib.ref_set_layer(which, _p(a))
