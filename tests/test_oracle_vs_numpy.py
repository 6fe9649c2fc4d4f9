"""Second opinion on the C oracle: an independent numpy float32 re-derivation of the reference
semantics (tests/np_reference.py) must agree bit for bit on keys / heights / variances / fused
layers / ray clean-up ("cross-validate the oracle two ways")."""
import numpy as np

import gem_b200
from gem_b200 import synth
import np_reference as npr
from oracle_lib import OracleMap

f32 = np.float32


def bits(a):
    return np.asarray(a, f32).view(np.uint32)


def _case(L, res, box, pos, fr, n=None):
    o = OracleMap(L, res, compat_box_filter=box)
    centre, start, _ = o.move(pos)
    sp = gem_b200.LaserSensorProcessor()
    f = gem_b200.make_frame(fr["T"], sp, base_z=0.2)
    xyzi = fr["xyzi"] if n is None else fr["xyzi"][:n]
    x, y, z = (xyzi[:, k].copy() for k in range(3))
    low0 = o.get_layer("lowest")
    ko = o.process_points(x, y, z, f)
    T32 = np.array(f.T[:], f32).reshape(4, 4)
    kn = npr.process_points(x, y, z, T32, f.rel_lower, f.rel_upper, L, res, centre, start, box, sp.min_radius,
                            sp.beam_angle, sp.beam_constant, T32[2, :3])
    key, geo, h, var, xt, yt = kn
    assert np.array_equal(ko[0], key)
    assert np.array_equal(bits(ko[1]), bits(var))
    assert np.array_equal(bits(ko[2]), bits(xt)) and np.array_equal(bits(ko[3]), bits(yt))
    assert np.array_equal(bits(ko[4]), bits(h))
    low = npr.lowest_update(low0, geo, h, var)
    assert np.array_equal(bits(o.get_layer("lowest").reshape(-1)), bits(low))
    return o, f, xyzi, key, h, var


def test_process_points_hdl64_even_and_odd():
    fr = synth.hdl64_frame(0)
    _case(200, 0.1, False, fr["position"], fr)
    _case(75, 0.2, False, fr["position"] + np.array([0.37, -0.2, 0]), fr)
    frc = synth.hdl64_frame(1, compat_axes=True)
    _case(120, 0.1, True, frc["position"], frc)


def test_fuse_small_cloud_matches_numpy():
    fr = synth.hdl64_frame(0)
    sub = {k: v for k, v in fr.items()}
    idx = np.arange(0, fr["xyzi"].shape[0], 9)[:9000]
    sub["xyzi"] = np.ascontiguousarray(fr["xyzi"][idx])
    sub["rgba"] = np.ascontiguousarray(fr["rgba"][idx])
    sub["rgba"][::7, 1] = 0
    o, f, xyzi, key, h, var = _case(60, 0.4, False, fr["position"], sub)
    R, G, B = (sub["rgba"][:, k].astype(np.int32) for k in range(3))
    st = [o.get_layer(n).reshape(-1) for n in ("elevation", "variance", "intensity", "color_r", "color_g", "color_b")]
    for rep in range(2):
        o.fuse_points(key, R, G, B, xyzi[:, 3], h, var)
        st = list(npr.fuse(*st, key, R, G, B, xyzi[:, 3], h, var))
    for name, ref in zip(("elevation", "variance", "intensity", "color_r", "color_g", "color_b"), st):
        got = o.get_layer(name).reshape(-1)
        if got.dtype.kind == "f":
            assert np.array_equal(bits(got), bits(ref)), name
        else:
            assert np.array_equal(got, ref), name


def test_raytracing_matches_python_dda():
    rng = np.random.default_rng(5)
    for L, start in ((20, (0, 0)), (21, (3, 17)), (32, (31, 5))):
        o = OracleMap(L, 0.1)
        o.m.contents.start[0], o.m.contents.start[1] = start
        o.m.contents.sensorZ = 0.8
        elev = np.where(rng.uniform(size=L * L) < 0.6, rng.uniform(-0.3, 1.2, L * L), -10).astype(f32)
        var = rng.uniform(1e-4, 0.01, L * L).astype(f32)
        trav = np.where(rng.uniform(size=L * L) < 0.5, rng.uniform(-0.5, 1.0, L * L), -10).astype(f32)
        low = np.where(rng.uniform(size=L * L) < 0.5, rng.uniform(-0.4, 0.6, L * L), 10).astype(f32)
        for name, arr in (("elevation", elev), ("variance", var), ("traver", trav), ("lowest", low)):
            o.set_layer(name, arr)
        want = npr.raytracing(elev, var, trav, low, L, start, f32(0.8))
        o.raytracing()
        got = o.get_layer("elevation").reshape(-1)
        assert np.array_equal(bits(got), bits(want)), L
        assert (got != elev).sum() > 0           # something was cleaned
        assert (o.get_layer("lowest") == 10).all()


def test_show_orthomosaic_and_visual_cloud_vs_numpy():
    """orc_show against a vectorised numpy statement of ElevationMap.cpp:97-125 (+ grid_map cell-centre positions)"""
    def laser_frame(T):
        return gem_b200.make_frame(T, gem_b200.LaserSensorProcessor())
    L, res = 96, 0.1
    o = OracleMap(L, res, compat_box_filter=False)
    for k in range(3):
        fr = synth.hdl64_frame(k, speed=9.0)
        o.move(fr["position"])
        o.add(fr["xyzi"], fr["rgba"], laser_frame(fr["T"]))
    f = o.map_feature()
    img, xyz, rgb = o.show()
    centre, start, _ = o.state()
    E, T = f["elevation"].reshape(L, L), f["traver"].reshape(L, L)
    valid = (E != -10) & (T != -10) & ~np.isnan(T)
    ix, iy = np.nonzero(valid.T)[1], np.nonzero(valid.T)[0]          # column-major visiting order: iy outer, ix inner
    ux, uy = (ix + L - start[0]) % L, (iy + L - start[1]) % L
    ref_img = np.zeros((L, L, 3), np.uint8)
    col = np.stack([f["color_b"].reshape(L, L)[ix, iy], f["color_g"].reshape(L, L)[ix, iy], f["color_r"].reshape(L, L)[ix, iy]], axis=1)
    ref_img[ux, uy] = col.astype(np.uint8)
    assert valid.sum() > 500 and np.array_equal(img, ref_img)
    half = 0.5 * L * np.float64(np.float32(res)) - 0.5 * np.float64(np.float32(res))
    px = (np.float64(centre[0]) + half - np.float64(np.float32(res)) * ux).astype(np.float32)
    py = (np.float64(centre[1]) + half - np.float64(np.float32(res)) * uy).astype(np.float32)
    assert np.array_equal(xyz[:, 0], px) and np.array_equal(xyz[:, 1], py) and np.array_equal(xyz[:, 2], E[ix, iy])
    assert np.array_equal(rgb, col[:, ::-1].astype(np.uint8))


def test_harvest_scrolled_out_vs_numpy():
    """orc_harvest (ElevationMapping.cpp:716-765) against a vectorised numpy statement, all eight direction cases"""
    L, res = 96, 0.1
    def laser_frame(T):
        return gem_b200.make_frame(T, gem_b200.LaserSensorProcessor())
    o = OracleMap(L, res, compat_box_filter=False)
    fr = synth.hdl64_frame(0)
    o.move(fr["position"])
    o.add(fr["xyzi"], fr["rgba"], laser_frame(fr["T"]))
    o.snapshot_shown()
    f, centre_p, start_p = o._prev
    p0 = np.array(fr["position"], np.float32)
    total = 0
    for dx, dy in [(0.7, 0.4), (-0.7, -0.4), (0.7, -0.4), (-0.7, 0.4), (0.5, 0.0), (-0.5, 0.0), (0.0, 0.5), (0.0, -0.5), (0.0, 0.0)]:
        o2 = OracleMap(L, res, compat_box_filter=False)
        o2.move(p0)
        cur, _, shift = o2.move(p0 + np.array([dx, dy, 0], np.float32))
        for grid_res in (0.0, 0.1):                       # float-derived and the node's double resolution
            rec, n = o.harvest_scrolled_out(cur, shift, grid_res=grid_res)
            r = np.float64(np.float32(res)) if grid_res == 0 else np.float64(grid_res)
            E, T = f["elevation"].reshape(L, L), f["traver"].reshape(L, L)
            shown = (E != -10) & (T != -10) & ~np.isnan(T) & (T >= 0)
            iy, ix = np.nonzero(shown.T)
            half = 0.5 * (L * r) - 0.5 * r
            x = np.float64(centre_p[0]) + half - r * ((ix + L - start_p[0]) % L)
            y = np.float64(centre_p[1]) + half - r * ((iy + L - start_p[1]) % L)
            hw = L * r / 2
            lox, hix, loy, hiy = cur[0] - hw, cur[0] + hw, cur[1] - hw, cur[1] + hw
            sx, sy = shift
            out = (((x < lox) | (y < loy)) & (sx > 0 and sy > 0)) | (((x > hix) | (y > hiy)) & (sx < 0 and sy < 0)) | \
                  (((x < lox) | (y > hiy)) & (sx > 0 and sy < 0)) | (((x > hix) | (y < loy)) & (sx < 0 and sy > 0)) | \
                  ((x < lox) & (sx > 0 and sy == 0)) | ((x > hix) & (sx < 0 and sy == 0)) | \
                  ((y < loy) & (sy > 0 and sx == 0)) | ((y > hiy) & (sy < 0 and sx == 0))
            ix, iy, x, y = ix[out], iy[out], x[out], y[out]
            assert n == ix.size
            if dx == 0 and dy == 0:
                assert n == 0
            assert np.array_equal(rec[:, 0], x.astype(np.float32)) and np.array_equal(rec[:, 1], y.astype(np.float32))
            assert np.array_equal(rec[:, 2], E[ix, iy]) and np.all(rec[:, 3] == 1)
            assert np.array_equal(rec[:, 5], f["variance"].reshape(L, L)[ix, iy])
            assert np.array_equal(bits(rec[:, 6]), bits(f["intensity"].reshape(L, L)[ix, iy]))
            assert np.array_equal(rec[:, 7], T[ix, iy])
            bgra = rec[:, 4].copy().view(np.uint32)
            assert np.array_equal(bgra & 255, f["color_b"].reshape(L, L)[ix, iy] & 255)
            assert np.array_equal((bgra >> 16) & 255, f["color_r"].reshape(L, L)[ix, iy] & 255)
            total += n
    assert total > 500
