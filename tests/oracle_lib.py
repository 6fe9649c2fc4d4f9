"""ctypes wrapper of the CPU oracle (oracle/libgem_oracle.so).  TEST INFRASTRUCTURE ONLY:
imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs; the product never imports it."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ODIR = os.path.join(ROOT, "oracle")
OLIB = os.path.join(ODIR, "libgem_oracle.so")


class OrcMap(C.Structure):
    _fields_ = [
        ("L", C.c_int), ("res", C.c_float), ("obstacle_threshold", C.c_float), ("mahalanobis", C.c_float),
        ("lowest", C.POINTER(C.c_float)), ("elevation", C.POINTER(C.c_float)), ("variance", C.POINTER(C.c_float)),
        ("intensity", C.POINTER(C.c_float)), ("traver", C.POINTER(C.c_float)),
        ("colorR", C.POINTER(C.c_int)), ("colorG", C.POINTER(C.c_int)), ("colorB", C.POINTER(C.c_int)),
        ("centre", C.c_float * 2), ("start", C.c_int * 2), ("sensorZ", C.c_float), ("compat_box_filter", C.c_int),
    ]


class OrcSensor(C.Structure):
    _fields_ = [("type", C.c_int), ("min_r", C.c_float), ("beam_a", C.c_float), ("beam_c", C.c_float),
                ("nf_a", C.c_double), ("nf_b", C.c_double), ("nf_c", C.c_double), ("nf_d", C.c_double),
                ("nf_e", C.c_double), ("lateral", C.c_double), ("cutoff_min", C.c_double), ("cutoff_max", C.c_double)]


_lib = None


def build():
    src = [os.path.join(ODIR, "gem_oracle.c"), os.path.join(ODIR, "gem_oracle.h")]
    if os.path.exists(OLIB) and all(os.path.getmtime(s) <= os.path.getmtime(OLIB) for s in src):
        return OLIB
    subprocess.run(["make", "-C", ODIR], check=True, stdout=subprocess.DEVNULL)
    return OLIB


def load():
    global _lib
    if _lib is not None:
        return _lib
    build()
    lib = C.CDLL(OLIB)
    P = C.c_void_p
    MP = C.POINTER(OrcMap)
    lib.orc_create.restype = MP
    lib.orc_create.argtypes = [C.c_int, C.c_float, C.c_float, C.c_float]
    lib.orc_destroy.argtypes = [MP]
    lib.orc_move.argtypes = [MP, P, P, P, P]
    lib.orc_points_to_index.restype = C.c_int
    lib.orc_points_to_index.argtypes = [MP, C.c_float, C.c_float, P]
    lib.orc_process_points.argtypes = [MP, C.c_int, P, P, P, P, C.c_double, C.c_double, C.POINTER(OrcSensor), P, P, P,
                                       P, P, P, P, P, P, P]
    lib.orc_fuse.argtypes = [MP, C.c_int, P, P, P, P, P, P, P]
    lib.orc_fuse_literal.argtypes = [MP, C.c_int, P, P, P, P, P, P, P]
    lib.orc_var_update.argtypes = [MP, C.c_float]
    lib.orc_map_feature.argtypes = [MP, P, P, P, P, P, P, P, P, P]
    lib.orc_raytracing.argtypes = [MP]
    lib.orc_optmove.argtypes = [MP, P, C.c_float, P]
    lib.orc_closeloop.argtypes = [MP, P, C.c_float]
    for fn in ("orc_sinf", "orc_cosf", "orc_acosf"):
        getattr(lib, fn).restype = C.c_float
        getattr(lib, fn).argtypes = [C.c_float]
    lib.orc_atan2f.restype = C.c_float
    lib.orc_atan2f.argtypes = [C.c_float, C.c_float]
    lib.orc_show.argtypes = [MP, C.c_double, P, P, P, P, P, P, P, P, C.POINTER(C.c_int)]
    lib.orc_harvest.argtypes = [C.c_int, C.c_double, P, P, P, P, P, P, P, P, P, P, P, P, C.POINTER(C.c_int)]
    lib.orc_colourise.argtypes = [P, C.c_int, P, P, P, C.c_int, C.c_int, C.c_int, P]
    lib.orc_clean_point_cloud.restype = C.c_int
    lib.orc_clean_point_cloud.argtypes = [C.POINTER(OrcSensor), C.c_int, P, P]
    lib.orc_add_points_mt.argtypes = [MP, C.c_int, P, P, P, C.c_double, C.c_double, C.POINTER(OrcSensor), P, C.c_int]
    lib.orc_transform_cloud.argtypes = [P, C.c_int, P]
    lib.orc_refuse_submaps.restype = C.c_int
    lib.orc_refuse_submaps.argtypes = [P, C.POINTER(C.c_int), P, C.POINTER(C.c_int), C.c_double, C.c_int]
    lib.orc_pool_create.restype = P
    lib.orc_pool_create.argtypes = [C.c_int]
    lib.orc_pool_destroy.argtypes = [P]
    lib.orc_add_points_pool.argtypes = [P, MP, C.c_int, P, P, P, C.c_double, C.c_double, C.POINTER(OrcSensor), P]
    _lib = lib
    return lib


def _p(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


def colourise(xyzi, T_camera, T_lidar, bgr):
    """oracle twin of gem_colourise_points: returns (xyzi with zeroed intensities, rgba)"""
    lib = load()
    xyzi = np.array(xyzi, np.float32, copy=True)
    tc = np.ascontiguousarray(T_camera, np.float64).reshape(-1)
    tl = np.ascontiguousarray(T_lidar, np.float64).reshape(-1)
    bgr = np.ascontiguousarray(bgr, np.uint8)
    rgba = np.zeros((xyzi.shape[0], 4), np.uint8)
    lib.orc_colourise(_p(xyzi), xyzi.shape[0], _p(tc), _p(tl), _p(bgr), bgr.shape[1], bgr.shape[0], 3 * bgr.shape[1], _p(rgba))
    return xyzi, rgba


def transform_cloud(pts, T):
    """oracle twin of gem_transform_cloud on an (n, 8) float32 array (copy)"""
    lib = load()
    pts = np.array(pts, np.float32, copy=True, order="C")
    T = np.ascontiguousarray(T, np.float32).reshape(-1)
    lib.orc_transform_cloud(_p(pts), pts.shape[0], _p(T))
    return pts


def refuse_submaps(new, old, res, compat=True):
    """oracle twin of gem_refuse_submaps: returns (new', old', fused)"""
    lib = load()
    new = np.array(new, np.float32, copy=True, order="C")
    old = np.array(old, np.float32, copy=True, order="C")
    nn, no = C.c_int(new.shape[0]), C.c_int(old.shape[0])
    fused = lib.orc_refuse_submaps(_p(new), C.byref(nn), _p(old), C.byref(no), float(res), 1 if compat else 0)
    return new[:nn.value], old[:no.value], fused


def sensor_from_frame(frame) -> OrcSensor:
    s = frame.sensor
    return OrcSensor(s.type, s.min_radius, s.beam_angle, s.beam_constant, s.normal_factor_a, s.normal_factor_b,
                     s.normal_factor_c, s.normal_factor_d, s.normal_factor_e, s.lateral_factor,
                     s.cutoff_min_depth, s.cutoff_max_depth)


class OracleMap:
    """Same method names as gem_b200.ElevationMap, executed by the CPU oracle."""

    def __init__(self, length, resolution, mahalanobis_threshold=2.5, obstacle_threshold=0.7, compat_box_filter=True):
        self.lib = load()
        self.m = self.lib.orc_create(int(length), float(resolution), float(mahalanobis_threshold),
                                     float(obstacle_threshold))
        self.m.contents.compat_box_filter = 1 if compat_box_filter else 0
        self.length = int(length)
        self.resolution = float(resolution)
        self.ncells = self.length * self.length
        self.shape = (self.length, self.length)

    def close(self):
        for pool in self.__dict__.pop("_pools", {}).values():
            self.lib.orc_pool_destroy(pool)
        if self.m:
            self.lib.orc_destroy(self.m)
            self.m = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def move(self, position):
        pos = np.asarray(position, np.float32)
        centre = np.zeros(2, np.float32)
        start = np.zeros(2, np.int32)
        shift = np.zeros(2, np.float32)
        self.lib.orc_move(self.m, _p(pos), _p(centre), _p(start), _p(shift))
        return centre, start, shift

    def process_points(self, x, y, z, frame):
        x = np.ascontiguousarray(x, np.float32)
        y = np.ascontiguousarray(y, np.float32)
        z = np.ascontiguousarray(z, np.float32)
        n = x.shape[0]
        key = np.empty(n, np.int32)
        var = np.empty(n, np.float32)
        xt = np.empty(n, np.float32)
        yt = np.empty(n, np.float32)
        zt = np.empty(n, np.float32)
        T = np.array(frame.T[:], np.float32)
        sJ = np.array(frame.sensor_jacobian[:], np.float32)
        rv = np.array(frame.rotation_variance[:], np.float32)
        cs = np.array(frame.C_SB_transpose[:], np.float32)
        pm = np.array(frame.P_mul_C_BM_transpose[:], np.float32)
        bs = np.array(frame.B_r_BS_skew[:], np.float32)
        sensor = sensor_from_frame(frame)
        self.lib.orc_process_points(self.m, n, _p(x), _p(y), _p(z), _p(T), frame.rel_lower, frame.rel_upper,
                                    C.byref(sensor), _p(sJ), _p(rv), _p(cs), _p(pm), _p(bs), _p(key), _p(var), _p(xt),
                                    _p(yt), _p(zt))
        return key, var, xt, yt, zt

    def fuse_points(self, index, R, G, B, intensity, height, var, literal=False):
        index = np.ascontiguousarray(index, np.int32)
        ci = lambda a: None if a is None else np.ascontiguousarray(a, np.int32)
        cf = lambda a: None if a is None else np.ascontiguousarray(a, np.float32)
        R, G, B, intensity, height, var = ci(R), ci(G), ci(B), cf(intensity), cf(height), cf(var)
        fn = self.lib.orc_fuse_literal if literal else self.lib.orc_fuse
        fn(self.m, index.shape[0], _p(index), _p(R), _p(G), _p(B), _p(intensity), _p(height), _p(var))

    def add(self, xyzi, rgba, frame, n=None):
        """process_points + fuse, i.e. what gem_add_points fuses on the device"""
        xyzi = np.asarray(xyzi, np.float32)
        if n is not None:
            xyzi = xyzi[:n]
            rgba = None if rgba is None else rgba[:n]
        xyzi, rgba = self.clean_point_cloud(xyzi, rgba, frame)
        key, var, xt, yt, zt = self.process_points(xyzi[:, 0], xyzi[:, 1], xyzi[:, 2], frame)
        if rgba is None:
            R = G = B = np.zeros(xyzi.shape[0], np.int32)
        else:
            R, G, B = (rgba[:, k].astype(np.int32) for k in range(3))
        self.fuse_points(key, R, G, B, xyzi[:, 3], zt, var)

    def clean_point_cloud(self, xyzi, rgba, frame):
        """SensorProcessorBase::process's cleanPointCloud (SPB.cpp:90): copies, like the reference (:83-87), and
        removes the points the sensor processor drops before Process_points"""
        xyzi = np.array(xyzi, np.float32, copy=True, order="C")
        rgba = None if rgba is None else np.array(rgba, np.uint8, copy=True, order="C")
        sensor = sensor_from_frame(frame)
        k = self.lib.orc_clean_point_cloud(C.byref(sensor), xyzi.shape[0], _p(xyzi), _p(rgba))
        return xyzi[:k], (None if rgba is None else rgba[:k])

    def add_mt(self, xyzi, rgba, frame, nthreads):
        xyzi, rgba = self.clean_point_cloud(xyzi, rgba, frame)
        xyzi = np.ascontiguousarray(xyzi, np.float32)
        rgba = None if rgba is None else np.ascontiguousarray(rgba, np.uint8)
        T = np.array(frame.T[:], np.float32)
        sJ = np.array(frame.sensor_jacobian[:], np.float32)
        sensor = sensor_from_frame(frame)
        self.lib.orc_add_points_mt(self.m, xyzi.shape[0], _p(xyzi), _p(rgba), _p(T), frame.rel_lower, frame.rel_upper,
                                   C.byref(sensor), _p(sJ), int(nthreads))

    def add_pool(self, xyzi, rgba, frame, nthreads):
        """pooled CPU baseline: persistent worker threads (created on first use, per thread count)"""
        xyzi, rgba = self.clean_point_cloud(xyzi, rgba, frame)
        pools = self.__dict__.setdefault("_pools", {})
        if nthreads not in pools:
            pools[nthreads] = self.lib.orc_pool_create(int(nthreads))
        T = np.array(frame.T[:], np.float32)
        sJ = np.array(frame.sensor_jacobian[:], np.float32)
        sensor = sensor_from_frame(frame)
        self.lib.orc_add_points_pool(pools[nthreads], self.m, xyzi.shape[0], _p(xyzi), _p(rgba), _p(T), frame.rel_lower,
                                     frame.rel_upper, C.byref(sensor), _p(sJ))

    def var_update(self, dv):
        self.lib.orc_var_update(self.m, float(dv))

    def map_feature(self):
        n = self.ncells
        out = {
            "elevation": np.empty(n, np.float32), "variance": np.empty(n, np.float32),
            "color_r": np.empty(n, np.int32), "color_g": np.empty(n, np.int32), "color_b": np.empty(n, np.int32),
            "rough": np.empty(n, np.float32), "slope": np.empty(n, np.float32), "traver": np.empty(n, np.float32),
            "intensity": np.empty(n, np.float32),
        }
        self.lib.orc_map_feature(self.m, _p(out["elevation"]), _p(out["variance"]), _p(out["color_r"]),
                                 _p(out["color_g"]), _p(out["color_b"]), _p(out["rough"]), _p(out["slope"]),
                                 _p(out["traver"]), _p(out["intensity"]))
        return out

    def compute_features(self):
        self.lib.orc_map_feature(self.m, None, None, None, None, None, None, None, None, None)

    def raytracing(self):
        self.lib.orc_raytracing(self.m)

    clean = raytracing

    def opt_move(self, opt_p, height_update):
        p = np.asarray(opt_p, np.float32)
        out = np.zeros(2, np.float32)
        self.lib.orc_optmove(self.m, _p(p), float(height_update), _p(out))
        return out

    def closeloop(self, update_position, height_update):
        p = np.asarray(update_position, np.float32)
        self.lib.orc_closeloop(self.m, _p(p), float(height_update))

    def get_layer(self, name):
        mm = self.m.contents
        src = {"elevation": (mm.elevation, np.float32), "variance": (mm.variance, np.float32),
               "intensity": (mm.intensity, np.float32), "color_r": (mm.colorR, np.int32),
               "color_g": (mm.colorG, np.int32), "color_b": (mm.colorB, np.int32), "traver": (mm.traver, np.float32),
               "lowest": (mm.lowest, np.float32)}[name]
        return np.ctypeslib.as_array(src[0], shape=(self.ncells,)).astype(src[1]).reshape(self.shape).copy()

    def set_layer(self, name, arr):
        mm = self.m.contents
        dst = {"elevation": mm.elevation, "variance": mm.variance, "intensity": mm.intensity, "color_r": mm.colorR,
               "color_g": mm.colorG, "color_b": mm.colorB, "traver": mm.traver, "lowest": mm.lowest}[name]
        view = np.ctypeslib.as_array(dst, shape=(self.ncells,))
        view[:] = np.asarray(arr).reshape(-1)

    def state(self):
        mm = self.m.contents
        return np.array(mm.centre[:], np.float32), np.array(mm.start[:], np.int32), float(mm.sensorZ)

    def points_to_index(self, px, py):
        st = C.c_int()
        g = self.lib.orc_points_to_index(self.m, float(np.float32(px)), float(np.float32(py)), C.byref(st))
        return g, st.value

    def snapshot_shown(self):
        """prevMap_ = map_.visualMap_ (ElevationMapping.cpp:422): Map_feature outputs + geometry of this frame"""
        centre, start, _ = self.state()
        self._prev = (self.map_feature(), np.array(centre, np.float32), np.array(start, np.int32))

    def harvest_scrolled_out(self, current_xy, shift_xy, grid_res=0.0):
        f, centre, start = self._prev
        L = self.length
        out = np.empty((L * L, 8), np.float32)
        cnt = C.c_int()
        cur = np.asarray(current_xy, np.float32)
        sh = np.asarray(shift_xy, np.float32)
        res = float(grid_res) if grid_res > 0 else float(np.float32(self.resolution))
        self.lib.orc_harvest(L, res, _p(centre), _p(start), _p(f["elevation"]), _p(f["variance"]), _p(f["traver"]),
                             _p(f["color_r"]), _p(f["color_g"]), _p(f["color_b"]), _p(f["intensity"]), _p(cur), _p(sh),
                             _p(out), C.byref(cnt))
        return out[:cnt.value].copy(), cnt.value

    def show(self, grid_res=0.0):
        """orthomosaic (L, L, 3) uint8 + visual cloud (xyz, rgb) of ElevationMap::show from map_feature() outputs"""
        f = self.map_feature()
        L = self.length
        img = np.empty((L, L, 3), np.uint8)
        xyz = np.empty((L * L, 3), np.float32)
        rgb = np.empty((L * L, 3), np.uint8)
        cnt = C.c_int()
        self.lib.orc_show(self.m, float(grid_res), _p(f["elevation"]), _p(f["traver"]), _p(f["color_r"]), _p(f["color_g"]),
                          _p(f["color_b"]), _p(img), _p(xyz), _p(rgb), C.byref(cnt))
        return img, xyz[:cnt.value].copy(), rgb[:cnt.value].copy()

    def export_layers(self):
        """ElevationMap::show's masking + grid_map column-major layout, from the oracle state
        (ElevationMap.cpp:97-110).  Requires map_feature() outputs."""
        f = self.map_feature()
        return export_from_feature(f, self.length)


def export_from_feature(f, L):
    mask = (f["elevation"] != -10) & (f["traver"] != -10) & ~np.isnan(f["traver"])
    out = {}
    for name in ["elevation", "variance", "rough", "slope", "traver", "color_r", "color_g", "color_b", "intensity"]:
        a = np.where(mask, f[name].astype(np.float32), np.float32(np.nan)).reshape(L, L)
        out[name] = np.asfortranarray(a)
    return out
