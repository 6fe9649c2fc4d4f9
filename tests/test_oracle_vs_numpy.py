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
