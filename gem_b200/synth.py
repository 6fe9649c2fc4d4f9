"""Seeded synthetic sensor frames for tests and benchmarks.

HDL-64E-shaped frames: 64 beams with elevation linspace(+2.0, -24.8) deg, 2083 azimuth steps
(133 312 rays, azimuth-major order like the spinning sensor), sensor 1.73 m above a ground
plane with +-2 cm noise, 40 axis-aligned boxes in +-25 m, range clip 0.9-120 m, misses dropped
(about 130 k returns).  intensity and r,g,b uniform in [1, 255] so the colour path fires.
Pose of frame f: yaw(theta_f) . trans(v * t), v = 10 m/s along +x at 10 Hz.
D435-shaped frames: 640x480 pinhole, fx = fy = 385, depth 0.2-3.25 m.

PRNG: numpy PCG64 seeded with 20240001 + frame index (scene: 20240000).
"""
from __future__ import annotations

import numpy as np

SCENE_SEED = 20240000
FRAME_SEED0 = 20240001
SENSOR_HEIGHT = 1.73


def make_scene(seed: int = SCENE_SEED, n_boxes: int = 40, extent: float = 25.0):
    rng = np.random.Generator(np.random.PCG64(seed))
    cx = rng.uniform(-extent, extent, n_boxes)
    cy = rng.uniform(-extent, extent, n_boxes)
    sx = rng.uniform(0.5, 4.0, n_boxes)
    sy = rng.uniform(0.5, 4.0, n_boxes)
    hz = rng.uniform(0.5, 3.0, n_boxes)
    lo = np.stack([cx - sx / 2, cy - sy / 2, np.zeros(n_boxes)], 1)
    hi = np.stack([cx + sx / 2, cy + sy / 2, hz], 1)
    # keep the corridor the robot drives along (y ~ 0) free so the sensor never sits in a box
    keep = ~((lo[:, 1] < 1.5) & (hi[:, 1] > -1.5))
    return lo[keep], hi[keep]


def _cast(origin, dirs, scene, rmin, rmax):
    """nearest hit distance of rays origin + t*dirs with ground z=0 and the boxes; inf = miss.
    Returns (t, is_ground)."""
    lo, hi = scene
    with np.errstate(divide="ignore", invalid="ignore"):
        t_ground = np.where(dirs[:, 2] < 0, -origin[2] / dirs[:, 2], np.inf)
        inv = 1.0 / dirs  # (n,3)
        t_best = np.full(dirs.shape[0], np.inf)
        for b in range(lo.shape[0]):
            t1 = (lo[b] - origin) * inv
            t2 = (hi[b] - origin) * inv
            tn = np.nanmax(np.minimum(t1, t2), axis=1)
            tf = np.nanmin(np.maximum(t1, t2), axis=1)
            hit = (tf >= tn) & (tf > 0)
            tb = np.where(hit, np.where(tn > 0, tn, tf), np.inf)
            t_best = np.minimum(t_best, tb)
    is_ground = t_ground <= t_best
    t = np.minimum(t_ground, t_best)
    t = np.where((t >= rmin) & (t <= rmax), t, np.inf)
    return t, is_ground


def pose_matrix(x, y, z, yaw):
    c, s = np.cos(yaw), np.sin(yaw)
    T = np.eye(4)
    T[0, 0], T[0, 1], T[1, 0], T[1, 1] = c, -s, s, c
    T[:3, 3] = (x, y, z)
    return T


def hdl64_pose(frame: int, speed: float = 10.0, rate_hz: float = 10.0):
    """(T map<-sensor 4x4 float64, track position [x, y, z])"""
    t = frame / rate_hz
    yaw = 0.05 * np.sin(0.1 * frame)
    x, y = speed * t, 0.0
    T = pose_matrix(x, y, SENSOR_HEIGHT, yaw)
    return T, np.array([x, y, SENSOR_HEIGHT])


def hdl64_frame(frame: int = 0, scene=None, compat_axes: bool = False, speed: float = 10.0):
    """Returns dict(xyzi (n,4) f32, rgba (n,4) u8, T (4,4) f64, position (3,))."""
    if scene is None:
        scene = make_scene()
    rng = np.random.Generator(np.random.PCG64(FRAME_SEED0 + frame))
    n_beams, n_az = 64, 2083
    elev = np.deg2rad(np.linspace(2.0, -24.8, n_beams))
    az = np.arange(n_az) * np.deg2rad(0.1728)
    AZ, EL = np.meshgrid(az, elev, indexing="ij")  # azimuth-major
    AZ, EL = AZ.reshape(-1), EL.reshape(-1)
    d_s = np.stack([np.cos(EL) * np.cos(AZ), np.cos(EL) * np.sin(AZ), np.sin(EL)], 1)
    T, pos = hdl64_pose(frame, speed)
    d_w = d_s @ T[:3, :3].T
    t, is_ground = _cast(T[:3, 3], d_w, scene, 0.9, 120.0)
    ok = np.isfinite(t)
    p_s = d_s[ok] * t[ok, None]
    noise = rng.uniform(-0.02, 0.02, p_s.shape[0])
    p_s[:, 2] += np.where(is_ground[ok], noise, 0.0)
    n = p_s.shape[0]
    inten = rng.integers(1, 256, n).astype(np.float32)
    rgba = rng.integers(1, 256, (n, 4)).astype(np.uint8)
    if compat_axes:
        # reference demo convention (README.md:133): sensor x left, y back.  Re-express the
        # same returns in that frame and fold the axis change into T.
        A = np.array([[0, -1, 0], [-1, 0, 0], [0, 0, 1]], float)  # p_std = A @ p_compat
        p_s = p_s @ A  # A is symmetric orthogonal => p_compat = A^T p_std = A p_std
        T = T.copy()
        T[:3, :3] = T[:3, :3] @ A
    xyzi = np.concatenate([p_s.astype(np.float32), inten[:, None]], 1).astype(np.float32)
    return {"xyzi": np.ascontiguousarray(xyzi), "rgba": np.ascontiguousarray(rgba), "T": T, "position": pos}


def d435_pose(frame: int, speed: float = 0.5, rate_hz: float = 30.0, height: float = 0.6, pitch_deg: float = 35.0):
    t = frame / rate_hz
    x = speed * t
    # optical frame: z forward, x right, y down; camera looks along +x_map pitched down
    p = np.deg2rad(pitch_deg)
    fwd = np.array([np.cos(p), 0.0, -np.sin(p)])
    right = np.array([0.0, -1.0, 0.0])
    down = np.cross(fwd, right)
    T = np.eye(4)
    T[:3, 0], T[:3, 1], T[:3, 2] = right, down, fwd
    T[:3, 3] = (x, 0.0, height)
    return T, np.array([x, 0.0, height])


def d435_frame(frame: int = 0, scene=None, raw: bool = True):
    if scene is None:
        scene = make_scene(SCENE_SEED + 7, n_boxes=60, extent=6.0)
    rng = np.random.Generator(np.random.PCG64(FRAME_SEED0 + 100000 + frame))
    W, H, fx, fy, cx, cy = 640, 480, 385.0, 385.0, 320.0, 240.0
    v, u = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
    u, v = u.reshape(-1), v.reshape(-1)
    d_s = np.stack([(u - cx) / fx, (v - cy) / fy, np.ones_like(u, float)], 1)  # z = 1 plane
    T, pos = d435_pose(frame)
    d_w = d_s @ T[:3, :3].T
    nrm = np.linalg.norm(d_w, axis=1)
    t, _ = _cast(T[:3, 3], d_w / nrm[:, None], scene, 0.0, 50.0)
    depth = t / nrm  # z in the optical frame
    # the full 640x480 image as the camera driver publishes it: pixels without a return carry NaN, returns
    # beyond the useful range keep their depth; dropping them is cleanPointCloud's job
    # (StructuredLightSensorProcessor.cpp:51-66), i.e. part of the path under test
    hit = np.isfinite(depth) & (depth <= 8.0)
    dd = np.where(hit, depth, np.nan)
    p_s = d_s * dd[:, None]
    p_s[:, 2] += rng.uniform(-0.002, 0.002, p_s.shape[0])
    if not raw:
        keep = np.isfinite(dd) & (dd >= 0.2) & (dd <= 3.25)
        p_s = p_s[keep]
    n = p_s.shape[0]
    inten = rng.integers(1, 256, n).astype(np.float32)
    rgba = rng.integers(1, 256, (n, 4)).astype(np.uint8)
    xyzi = np.concatenate([p_s.astype(np.float32), inten[:, None]], 1).astype(np.float32)
    return {"xyzi": np.ascontiguousarray(xyzi), "rgba": np.ascontiguousarray(rgba), "T": T, "position": pos}


def random_cloud(n: int, seed: int, extent: float = 12.0, zmin: float = -1.0, zmax: float = 1.5,
                 zero_colour_frac: float = 0.1, dup_frac: float = 0.3):
    """Uniform random cloud with many same-cell collisions (for order-dependence tests)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    xyz = np.stack([rng.uniform(-extent, extent, n), rng.uniform(-extent, extent, n), rng.uniform(zmin, zmax, n)], 1)
    ndup = int(n * dup_frac)
    if ndup > 0 and n > 1:
        src = rng.integers(0, n, ndup)
        dst = rng.integers(0, n, ndup)
        xyz[dst, :2] = xyz[src, :2] + rng.uniform(-0.01, 0.01, (ndup, 2))
    inten = rng.integers(0, 256, n).astype(np.float32)
    rgba = rng.integers(0, 256, (n, 4)).astype(np.uint8)
    zc = rng.uniform(size=n) < zero_colour_frac
    rgba[zc, rng.integers(0, 3)] = 0
    xyzi = np.concatenate([xyz.astype(np.float32), inten[:, None]], 1).astype(np.float32)
    return {"xyzi": np.ascontiguousarray(xyzi), "rgba": np.ascontiguousarray(rgba)}
