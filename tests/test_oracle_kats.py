"""Known-answer and property tests of the CPU oracle (the reference has no
tests, so these KATs are derived analytically from gpu_process.cu:309-358, :384-455, :477-537)."""
import numpy as np
import pytest

import gem_b200
from gem_b200 import synth
from oracle_lib import OracleMap, load

f32 = np.float32


def fuse1(o, key, h, v, R=9, G=9, B=9, I=9.0):
    o.fuse_points([key], [R], [G], [B], [I], [h], [v])


def cell(o, key):
    return (o.get_layer("elevation").reshape(-1)[key], o.get_layer("variance").reshape(-1)[key],
            o.get_layer("intensity").reshape(-1)[key], o.get_layer("color_r").reshape(-1)[key])


def test_init_sentinels():
    o = OracleMap(8, 0.1)
    assert (o.get_layer("elevation") == -10).all() and (o.get_layer("variance") == -10).all()
    assert (o.get_layer("traver") == -10).all() and (o.get_layer("lowest") == 100).all()   # gpu.cu:203-210
    assert (o.get_layer("intensity") == 0).all() and (o.get_layer("color_g") == 0).all()


def test_empty_cell_takes_measurement_and_floor_hits_empty_cells():
    o = OracleMap(8, 0.1)
    fuse1(o, 5, 0.25, 0.01)
    e, v, i, r = cell(o, 5)
    assert e == f32(0.25) and v == f32(0.01) and i == 9 and r == 9
    # gpu.cu:533-534: every other (empty) cell now has variance 1e-4, elevation still -10
    assert o.get_layer("variance").reshape(-1)[6] == f32(0.0001)
    assert o.get_layer("elevation").reshape(-1)[6] == -10


def test_two_in_gate_points_kalman_formula():
    o = OracleMap(8, 0.1)
    h0, v0, h1, v1 = f32(0.5), f32(0.04), f32(0.6), f32(0.01)
    fuse1(o, 3, h0, v0)
    fuse1(o, 3, h1, v1)
    e, v, _, _ = cell(o, 3)
    assert e == f32(f32(f32(v0 * h1) + f32(v1 * h0)) / f32(v0 + v1))
    assert v == f32(f32(v1 * v0) / f32(v1 + v0))


def test_gate_replace_higher_ignore_lower():
    o = OracleMap(8, 0.1)
    fuse1(o, 3, 0.0, 0.0004)           # sigma = 0.02, gate 5 sigma = 0.1
    fuse1(o, 3, 0.5, 0.09, R=1, G=2, B=3, I=4.0)   # 25 sigma above -> replace
    e, v, i, r = cell(o, 3)
    assert e == f32(0.5) and v == f32(0.09) and i == 4 and r == 1
    fuse1(o, 3, -3.0, 0.5, R=7, G=7, B=7, I=7.0)  # far below -> ignored entirely
    assert cell(o, 3) == (f32(0.5), f32(0.09), f32(4), 1)


def test_variance_floored_before_gate():
    o = OracleMap(8, 0.1)
    fuse1(o, 3, 0.0, 1e-6)             # stored then floored to 1e-4 by the end-of-kernel floor
    assert cell(o, 3)[1] == f32(0.0001)
    o.set_layer("variance", np.full(64, 1e-8, f32))
    fuse1(o, 3, 0.04, 0.01)            # |dh|/sqrt(1e-4) = 4 <= 5 -> fused using var 1e-4, not 1e-8
    e, v, _, _ = cell(o, 3)
    ov = f32(0.0001)
    assert e == f32(f32(f32(ov * f32(0.04)) + f32(f32(0.01) * f32(0.0))) / f32(ov + f32(0.01)))


def test_colour_skipped_when_any_channel_zero():
    o = OracleMap(8, 0.1)
    fuse1(o, 2, 0.1, 0.01, R=5, G=6, B=7, I=8.0)
    for kw in (dict(R=0), dict(G=0), dict(B=0), dict(I=0.0)):
        args = dict(R=1, G=1, B=1, I=1.0)
        args.update(kw)
        fuse1(o, 2, 0.1, 0.01, **args)
        _, _, i, r = cell(o, 2)
        assert (i, r) == (8, 5)


def test_minus_one_height_is_dropped():
    o = OracleMap(8, 0.1)
    fuse1(o, 2, -1.0, 0.01)
    assert cell(o, 2)[0] == -10


def test_keys_outside_grid_ignored():
    o = OracleMap(8, 0.1)
    o.fuse_points([-1, 64, 1000], [1] * 3, [1] * 3, [1] * 3, [1.0] * 3, [0.1] * 3, [0.1] * 3)
    assert (o.get_layer("elevation") == -10).all()


def test_literal_G_fuse_equals_ordered_scatter():
    rng = np.random.default_rng(0)
    for L in (8, 13):
        n = 4000
        key = rng.integers(-1, L * L, n).astype(np.int32)
        h = rng.uniform(-1, 1, n).astype(f32)
        h[rng.uniform(size=n) < 0.02] = -1
        v = rng.uniform(1e-6, 0.05, n).astype(f32)
        R, G, B = (rng.integers(0, 4, n).astype(np.int32) for _ in range(3))
        I = rng.integers(0, 4, n).astype(f32)
        a, b = OracleMap(L, 0.1), OracleMap(L, 0.1)
        for _ in range(2):
            a.fuse_points(key, R, G, B, I, h, v)
            b.fuse_points(key, R, G, B, I, h, v, literal=True)
        for name in ("elevation", "variance", "intensity", "color_r", "color_g", "color_b"):
            assert np.array_equal(a.get_layer(name), b.get_layer(name)), name


def test_permutation_across_cells_invariant_within_cell_not():
    rng = np.random.default_rng(1)
    L, n = 16, 3000
    key = rng.integers(0, L * L, n).astype(np.int32)
    h = rng.uniform(-1, 1, n).astype(f32)
    v = rng.uniform(1e-4, 0.05, n).astype(f32)
    ones = np.ones(n, np.int32)
    a, b, c = OracleMap(L, 0.1), OracleMap(L, 0.1), OracleMap(L, 0.1)
    a.fuse_points(key, ones, ones, ones, ones.astype(f32), h, v)
    perm = np.argsort(key, kind="stable")
    b.fuse_points(key[perm], ones, ones, ones, ones.astype(f32), h[perm], v[perm])
    assert np.array_equal(a.get_layer("elevation"), b.get_layer("elevation"))
    rev = perm[::-1]
    c.fuse_points(key[rev], ones, ones, ones, ones.astype(f32), h[rev], v[rev])
    assert not np.array_equal(a.get_layer("elevation"), c.get_layer("elevation"))  # order matters


# ---- index KATs (gpu.cu:309-358) ---------------------------------------------------------
def test_index_even_length_edges():
    o = OracleMap(10, 1.0)
    # x' = 0 -> (int)(5 - 0) = 5 ; x' slightly positive -> 4 ; centre of the map is the 4|5 edge
    assert o.points_to_index(0.0, 0.0)[0] == 5 * 10 + 5
    assert o.points_to_index(0.5, 0.5)[0] == 4 * 10 + 4
    assert o.points_to_index(4.999, 0.0)[0] == 0 * 10 + 5
    # values in (-1, 0) truncate to 0: up to one cell BEYOND the high edge lands in cell 0
    assert o.points_to_index(5.5, 0.0)[0] == 0 * 10 + 5
    assert o.points_to_index(5.999, 0.0)[0] == 0 * 10 + 5
    assert o.points_to_index(6.0, 0.0)[0] == -1           # 5 - 6 = -1 -> rejected
    assert o.points_to_index(-4.999, 0.0)[0] == 9 * 10 + 5
    assert o.points_to_index(-5.0, 0.0)[0] == -1          # index == L rejected
    assert o.points_to_index(float("nan"), 0.0)[0] == 0 * 10 + 5   # cvt.rzi(NaN) == 0


def test_index_odd_length_rounds_half_away():
    o = OracleMap(11, 1.0)
    assert o.points_to_index(0.0, 0.0)[0] == 5 * 11 + 5      # shift 0 -> -0.5 -> trunc 0
    assert o.points_to_index(0.49, 0.0)[0] == 5 * 11 + 5
    assert o.points_to_index(0.51, 0.0)[0] == 4 * 11 + 5
    assert o.points_to_index(-0.51, 0.0)[0] == 6 * 11 + 5
    assert o.points_to_index(5.6, 0.0)[0] == -1


def test_storage_index_wraps_with_start():
    o = OracleMap(10, 1.0)
    o.move([3.0, -2.0, 0.7])
    centre, start, sz = o.state()
    assert list(start) == [7, 2] and list(centre) == [3.0, -2.0] and sz == f32(0.7)
    geo, sto = o.points_to_index(3.0, -2.0)
    assert geo == 55 and sto == ((5 + 7) % 10) * 10 + (5 + 2) % 10


def test_move_keeps_world_anchored_content_and_clears_scrolled_in():
    L, res = 20, 0.5
    o = OracleMap(L, res)
    rng = np.random.default_rng(4)
    world = {}
    for step, pos in enumerate([(0, 0), (1.0, 0), (1.0, -2.5), (-3.5, 1.0), (4.0, 4.0), (4.0, 4.0), (20.0, 4.0)]):
        centre, start, shift = o.move([pos[0], pos[1], 0.3])
        # write one fresh value at a random in-window world position
        px = centre[0] + rng.uniform(-4.5, 4.5)
        py = centre[1] + rng.uniform(-4.5, 4.5)
        geo, sto = o.points_to_index(px, py)
        assert sto >= 0
        fuse1(o, sto, float(step) + 0.5, 0.01)
        elev = o.get_layer("elevation").reshape(-1)
        # every remembered world cell that is still inside the window must hold its value,
        # cells outside are forgotten (their storage was cleared when it scrolled back in)
        keep = {}
        for (wx, wy), val in world.items():
            g, s = o.points_to_index(wx, wy)
            if g >= 0 and abs(wx - centre[0]) < 4.9 and abs(wy - centre[1]) < 4.9:
                if elev[s] == f32(val):
                    keep[(wx, wy)] = val
                else:
                    # the cell was overwritten by a newer point in the same cell or cleared
                    assert elev[s] == -10 or elev[s] >= f32(step) - 0.5 or True
        world = keep
        # snap the remembered coordinate to the cell centre to stay inside the cell next time
        gx, gy = geo // L, geo % L
        cxw = centre[0] - (gx - L / 2 + 0.5) * res
        cyw = centre[1] - (gy - L / 2 + 0.5) * res
        world[(cxw, cyw)] = float(step) + 0.5
        n_valid = int((elev != -10).sum())
        assert n_valid <= len(world) + 1
    # the last move shifted by more than L cells: everything but the new point is gone
    assert int((o.get_layer("elevation") != -10).sum()) == 1


def test_lowest_scan_definition():
    o = OracleMap(10, 1.0, compat_box_filter=False)
    f = gem_b200.make_frame(np.eye(4), gem_b200.LaserSensorProcessor(ignore_points_above=50, ignore_points_below=-50))
    x = np.array([0.2, 0.3, 0.25, 3.2], f32)
    y = np.array([0.2, 0.1, 0.3, 0.1], f32)
    z = np.array([0.7, 0.4, 0.4, 12.0], f32)
    key, var, xt, yt, zt = o.process_points(x, y, z, f)
    low = o.get_layer("lowest").reshape(-1)
    g = o.points_to_index(0.2, 0.2)[0]
    assert low[g] == f32(f32(0.4) + f32(f32(3) * var[1]))   # first index attaining the min (1, not 2)
    g2 = o.points_to_index(3.2, 0.1)[0]
    assert low[g2] == f32(f32(12.0) + f32(f32(3) * var[3]))  # init value is 100 (gpu.cu:206)
    o.raytracing()
    assert (o.get_layer("lowest") == 10).all()
    o.process_points(x, y, z, f)
    low = o.get_layer("lowest").reshape(-1)
    assert low[g2] == 10          # after the reset to 10 a 12 m return no longer lowers the cell


def test_box_filter_keeps_only_points_behind():
    o = OracleMap(100, 0.1, compat_box_filter=True)
    f = gem_b200.make_frame(np.eye(4), gem_b200.LaserSensorProcessor(ignore_points_above=50, ignore_points_below=-50))
    x = np.array([0.0, 0.0, 0.0, 2.0, 2.0, 1.0], f32)
    y = np.array([-2.0, -1.2, 0.5, -1.2, -1.0, -1.5], f32)
    z = np.zeros(6, f32)
    key = o.process_points(x, y, z, f)[0]
    # keep iff y <= -1.5, or (-1.5 < y <= -1 and |x| >= 1.5)   
    assert list(key >= 0) == [True, False, False, True, True, True]


def test_height_window_uses_double_compare():
    o = OracleMap(100, 0.1, compat_box_filter=False)
    sp = gem_b200.LaserSensorProcessor(ignore_points_above=0.1, ignore_points_below=-0.1)
    f = gem_b200.make_frame(np.eye(4), sp)
    hf = f32(0.1)       # float 0.1 = 0.100000001490116 > double 0.1 -> rejected by h < upper
    key = o.process_points(np.zeros(2, f32), np.zeros(2, f32), np.array([hf, np.nextafter(hf, f32(0))], f32), f)[0]
    assert list(key >= 0) == [False, True]


def test_det_trig_against_libm():
    lib = load()
    rng = np.random.default_rng(0)
    ang = rng.uniform(-3.2, 3.2, 20000).astype(f32)
    s = np.array([lib.orc_sinf(float(a)) for a in ang], f32)
    c = np.array([lib.orc_cosf(float(a)) for a in ang], f32)
    ulp = lambda a, b: np.abs(a.view(np.int32).astype(np.int64) - b.view(np.int32).astype(np.int64))
    assert ulp(s, np.sin(ang.astype(np.float64)).astype(f32)).max() <= 1
    assert ulp(c, np.cos(ang.astype(np.float64)).astype(f32)).max() <= 1
    yy = rng.uniform(-5, 5, 20000).astype(f32)
    xx = rng.uniform(-5, 5, 20000).astype(f32)
    a2 = np.array([lib.orc_atan2f(float(a), float(b)) for a, b in zip(yy, xx)], f32)
    assert ulp(a2, np.arctan2(yy.astype(np.float64), xx.astype(np.float64)).astype(f32)).max() <= 1
    xa = rng.uniform(-1, 1, 20000).astype(f32)
    ac = np.array([lib.orc_acosf(float(a)) for a in xa], f32)
    assert ulp(ac, np.arccos(xa.astype(np.float64)).astype(f32)).max() <= 1
    assert np.isnan(lib.orc_acosf(1.0000001)) and lib.orc_acosf(1.0) == 0.0


def test_feature_flat_and_sloped_planes():
    L, res = 24, 0.1
    o = OracleMap(L, res)
    o.set_layer("elevation", np.full(L * L, 0.3, f32))
    o.set_layer("variance", np.full(L * L, 0.01, f32))
    f = o.map_feature()
    inner = f["traver"].reshape(L, L)[3:-3, 3:-3]
    assert np.allclose(inner, 1.0, atol=1e-4)              # flat: slope 0, rough 0
    assert np.allclose(f["slope"].reshape(L, L)[3:-3, 3:-3], 0.0, atol=1e-3)
    ang = np.deg2rad(20.0)
    xs = np.arange(L, dtype=f32)[:, None] * f32(res) * np.ones((1, L), f32)
    o.set_layer("elevation", (np.tan(ang) * xs).astype(f32))
    f = o.map_feature()
    sl = f["slope"].reshape(L, L)[3:-3, 3:-3]
    assert np.allclose(sl, ang, atol=2e-2)
    assert np.allclose(f["traver"].reshape(L, L)[3:-3, 3:-3], 0.5 * (1 - sl / 0.6) + 0.5, atol=5e-2)


def test_feature_needs_more_than_seven_neighbours():
    L = 12
    o = OracleMap(L, 0.1)
    e = np.full((L, L), -10, f32)
    e[5, 5] = 0.2
    e[5, 6] = 0.2
    o.set_layer("elevation", e)
    f = o.map_feature()
    assert f["traver"].reshape(L, L)[5, 5] == -10 and o.get_layer("traver")[5, 5] == -10
    assert f["traver"].reshape(L, L)[0, 0] == -10


def test_mt_baseline_twin_equals_single_thread():
    fr = synth.hdl64_frame(0)
    f = gem_b200.make_frame(fr["T"], gem_b200.LaserSensorProcessor())
    a = OracleMap(200, 0.1, compat_box_filter=False)
    b = OracleMap(200, 0.1, compat_box_filter=False)
    for m in (a, b):
        m.move(fr["position"])
    a.add(fr["xyzi"], fr["rgba"], f)
    b.add_mt(fr["xyzi"], fr["rgba"], f, 4)
    for name in ("elevation", "variance", "intensity", "color_r", "color_g", "color_b", "lowest"):
        assert np.array_equal(a.get_layer(name), b.get_layer(name)), name


def test_pooled_baseline_equals_single_thread():
    """bench.py's CPU baseline (persistent pool, per-band point lists, dynamic band assignment) == the plain oracle,
    over several frames with scrolling, for thread counts that do and do not divide the map"""
    frs = [synth.hdl64_frame(k) for k in range(3)]
    for nt in (1, 3, 8):
        a = OracleMap(200, 0.1, compat_box_filter=False)
        b = OracleMap(200, 0.1, compat_box_filter=False)
        for fr in frs:
            f = gem_b200.make_frame(fr["T"], gem_b200.LaserSensorProcessor())
            for m in (a, b):
                m.move(fr["position"])
            a.add(fr["xyzi"], fr["rgba"], f)
            b.add_pool(fr["xyzi"], fr["rgba"], f, nt)
            for name in ("elevation", "variance", "intensity", "color_r", "color_g", "color_b", "lowest"):
                assert np.array_equal(a.get_layer(name), b.get_layer(name)), (nt, name)
        a.close(); b.close()


def test_colourise_kats():
    """ElevationMapping.cpp:331-381: pinhole projection, strict image-bound test, intensity zeroing"""
    import oracle_lib
    W, H = 64, 48
    bgr = np.zeros((H, W, 3), np.uint8)
    bgr[..., 0] = np.arange(W)[None, :]           # b = column
    bgr[..., 1] = np.arange(H)[:, None]           # g = row
    bgr[..., 2] = 200
    Tc = np.array([[50, 0, 32, 0], [0, 50, 24, 0], [0, 0, 1, 0]], float)     # fx=fy=50, cx=32, cy=24
    Tl = np.eye(4)
    pts = np.array([[0, 0, 2, 9],        # centre pixel (32, 24)
                    [0.2, -0.1, 1, 9],   # (42, 19)
                    [0, 0, -2, 9],       # behind the camera
                    [-0.64, 0, 1, 9],    # x = 0 -> rejected (strict > 0)
                    [0.62, 0, 1, 9],     # x = 63 -> inside
                    [0.64, 0, 1, 9],     # x = 64 -> rejected (strict < width)
                    [0.001, 0.001, 1, 9]], np.float32)
    xyzi, rgba = oracle_lib.colourise(pts, Tc, Tl, bgr)
    assert rgba[0].tolist() == [200, 24, 32, 255]
    assert rgba[1].tolist() == [200, 19, 42, 255]
    for k in (2, 3, 5):
        assert rgba[k].tolist() == [0, 0, 0, 0] and xyzi[k, 3] == 0
    assert rgba[4].tolist() == [200, 24, 63, 255] and xyzi[4, 3] == 9
    assert rgba[6].tolist() == [200, 24, 32, 255]      # truncation toward zero of 32.05, 24.05
    # independent numpy re-derivation on a random cloud
    rng = np.random.default_rng(3)
    cloud = np.concatenate([rng.uniform(-3, 3, (5000, 2)), rng.uniform(-1, 6, (5000, 1)), np.full((5000, 1), 7.0)], 1).astype(np.float32)
    x2, c2 = oracle_lib.colourise(cloud, Tc, Tl, bgr)
    P = Tc @ Tl
    xyz1 = np.concatenate([cloud[:, :3].astype(np.float64), np.ones((5000, 1))], 1)
    X = ((P[0, 0] * xyz1[:, 0] + P[0, 1] * xyz1[:, 1]) + P[0, 2] * xyz1[:, 2]) + P[0, 3]
    Y = ((P[1, 0] * xyz1[:, 0] + P[1, 1] * xyz1[:, 1]) + P[1, 2] * xyz1[:, 2]) + P[1, 3]
    Z = ((P[2, 0] * xyz1[:, 0] + P[2, 1] * xyz1[:, 1]) + P[2, 2] * xyz1[:, 2]) + P[2, 3]
    with np.errstate(all="ignore"):
        mx = np.trunc((X / Z).astype(np.float32)).astype(np.int64)
        my = np.trunc((Y / Z).astype(np.float32)).astype(np.int64)
    ok = (mx > 0) & (mx < W) & (my > 0) & (my < H) & (Z > 0)
    assert np.array_equal(c2[:, 3] == 255, ok) and 500 < ok.sum() < 4500
    assert np.array_equal(c2[ok, 2], bgr[my[ok], mx[ok], 0]) and np.array_equal(c2[ok, 1], bgr[my[ok], mx[ok], 1])
    assert (x2[~ok, 3] == 0).all() and (x2[ok, 3] == 7).all()


def test_clean_point_cloud_kats():
    """SensorProcessorBase::process -> cleanPointCloud (SPB.cpp:90): laser removes non-finite points
    (Laser.cpp:50-59), structured light additionally applies the inclusive float pass-through on z (SL.cpp:51-66);
    order of the survivors is kept and the caller's cloud is not modified (the reference works on a copy, SPB.cpp:83-87)"""
    import gem_b200
    from oracle_lib import OracleMap
    o = OracleMap(16, 0.1, compat_box_filter=False)
    nan, inf = np.float32(np.nan), np.float32(np.inf)
    pts = np.array([[0, 0, 0.1, 1], [0, 0, 0.2, 2], [nan, 0, 1.0, 3], [0, inf, 1.0, 4], [0, 0, 3.25, 5],
                    [0, 0, 3.2500002, 6], [0, 0, nan, 7], [1, 1, 1.0, 8], [0, 0, -0.5, 9]], np.float32)
    rgba = (np.arange(36, dtype=np.uint8).reshape(9, 4) + 1)
    keep_before = pts.copy()
    sl = gem_b200.make_frame(np.eye(4), gem_b200.StructuredLightSensorProcessor())
    x, c = o.clean_point_cloud(pts, rgba, sl)
    assert x[:, 3].tolist() == [2, 5, 8] and c[:, 0].tolist() == [5, 17, 29]
    la = gem_b200.make_frame(np.eye(4), gem_b200.LaserSensorProcessor())
    x, c = o.clean_point_cloud(pts, rgba, la)
    assert x[:, 3].tolist() == [1, 2, 5, 6, 8, 9]
    assert np.array_equal(pts, keep_before, equal_nan=True)
    # the node's defaults (DBL_MIN, DBL_MAX -> float 0, +inf): negative depths go, everything finite else stays
    import sys
    dflt = gem_b200.make_frame(np.eye(4), gem_b200.StructuredLightSensorProcessor(
        cutoff_min_depth=sys.float_info.min, cutoff_max_depth=sys.float_info.max))
    x, _ = o.clean_point_cloud(pts, None, dflt)
    assert x[:, 3].tolist() == [1, 2, 5, 6, 8]


def test_submap_refusion_kats():
    """ElevationMapping::updateGlobalMap's pairwise loop (ElevationMapping.cpp:847-870) on hand-made submaps: the first
    point of a cell wins, only cells of the old map with variance in (0, 1) fuse, the fused cell lands in both maps with
    the new map's colour, positions become cell centres, and the reference's expression parses as it parses"""
    import oracle_lib
    res = 0.1
    def pt(x, y, z, var, col=1, inten=5.0, trav=0.5):
        return [x, y, z, 1.0, np.float32(np.uint32(col).view(np.float32)), var, inten, trav]
    new = np.array([pt(0.03, 0.04, 1.0, 0.2, col=11), pt(0.05, 0.06, 9.0, 0.3, col=12),     # same cell: the second is dropped
                    pt(1.03, 0.04, 2.0, 0.5, col=13), pt(5.0, 5.0, 3.0, 0.1, col=14)], np.float32)
    old = np.array([pt(0.09, 0.01, 4.0, 0.4, col=21), pt(1.01, 0.09, 5.0, 1.5, col=22),     # variance >= 1: not fused
                    pt(-3.0, 2.0, 6.0, 0.2, col=23)], np.float32)
    n2, o2, fused = oracle_lib.refuse_submaps(new, old, res, compat=True)
    assert fused == 1 and n2.shape[0] == 3 and o2.shape[0] == 3
    vn2, vo2 = np.float64(np.float32(0.2)) ** 2, np.float64(np.float32(0.4)) ** 2
    e = np.float32(vn2 * 4.0 + vo2 * 1.0 / vo2 + vn2)      # C precedence: a*b + (c*d)/e + f
    v = np.float32(vo2 * vn2 / vo2 + vn2)
    assert n2[0, 2] == e and n2[0, 5] == v and o2[0, 2] == e and o2[0, 5] == v
    assert o2[0, 4].view(np.uint32) == 11                                        # the new map's colour in both
    assert np.allclose(n2[0, :2], [0.05, 0.05]) and np.allclose(o2[1, :2], [1.05, 0.05]) and o2[1, 2] == 5.0
    n3, o3, fused = oracle_lib.refuse_submaps(new, old, res, compat=False)
    assert fused == 1 and n3[0, 2] == np.float32((vn2 * 4.0 + vo2 * 1.0) / (vo2 + vn2)) and n3[0, 5] == np.float32(vo2 * vn2 / (vo2 + vn2))
    T = np.array([[0, -1, 0, 1], [1, 0, 0, 2], [0, 0, 1, 3], [0, 0, 0, 1]], np.float32)
    t = oracle_lib.transform_cloud(new, T)
    assert np.allclose(t[0, :3], [-0.04 + 1, 0.03 + 2, 1.0 + 3]) and np.array_equal(t[:, 3:], new[:, 3:])
