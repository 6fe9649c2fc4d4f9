"""Independent numpy float32 re-derivation of the reference semantics (second opinion on the C
oracle; SURVEY.md 8c asks for it).  Written from gpu_process.cu, not from oracle/gem_oracle.c.
Slow Python loops: use small inputs only."""
from __future__ import annotations

import math

import numpy as np

f32 = np.float32


def _rz(v):
    """cvt.rzi.s32: truncate toward zero, saturate, NaN -> 0 (vectorised, float64/float32 in)"""
    v = np.asarray(v, dtype=np.float64)
    out = np.where(np.isnan(v), 0.0, np.trunc(v))
    out = np.clip(out, -2147483648.0, 2147483647.0)
    return out.astype(np.int64).astype(np.int32)


def points_to_index(px, py, L, res, centre, start):
    """gpu.cu:309-358 -> (geo linear or -1, storage linear or -1)"""
    px = np.asarray(px, f32)
    py = np.asarray(py, f32)
    sx = (px - f32(centre[0])).astype(f32)
    sy = (py - f32(centre[1])).astype(f32)
    with np.errstate(all="ignore"):
        if L % 2 == 0:
            ix = _rz((f32(L // 2) - (sx / f32(res)).astype(f32)).astype(f32))
            iy = _rz((f32(L // 2) - (sy / f32(res)).astype(f32)).astype(f32))
        else:
            ix = L // 2 - _rz((sx / f32(res)).astype(f32).astype(np.float64) + 0.5 * np.where(sx > 0, 1, -1))
            iy = L // 2 - _rz((sy / f32(res)).astype(f32).astype(np.float64) + 0.5 * np.where(sy > 0, 1, -1))
    ok = (ix >= 0) & (ix < L) & (iy >= 0) & (iy < L)
    geo = np.where(ok, ix * L + iy, -1)
    sto = np.where(ok, ((ix + start[0]) % L) * L + (iy + start[1]) % L, -1)
    return geo.astype(np.int32), sto.astype(np.int32)


def process_points(x, y, z, T, rel_lower, rel_upper, L, res, centre, start, box_filter, min_r, beam_a, beam_c, sJ):
    """gpu.cu:384-455 with rotationVariance == 0, laser model.  Returns key, geo, h, var, xt, yt."""
    x = np.asarray(x, f32); y = np.asarray(y, f32); z = np.asarray(z, f32)
    T = np.asarray(T, f32).reshape(4, 4)
    mul = lambda a, b: (a * b).astype(f32)
    add = lambda a, b: (a + b).astype(f32)
    with np.errstate(all="ignore"):
        h = add(add(add(mul(T[2, 0], x), mul(T[2, 1], y)), mul(T[2, 2], z)), T[2, 3])
        flag = np.zeros(x.shape, bool)
        if box_filter:
            flag = ((x > -1.5) & (x < 1.5) & (y > -1.5) & (y < 1.5)) | ((y > -1) & (y < 1)) | (y > 0)
        acc = (h.astype(np.float64) > rel_lower) & (h.astype(np.float64) < rel_upper) & ~flag
        xt = add(add(add(mul(T[0, 0], x), mul(T[0, 1], y)), mul(T[0, 2], z)), T[0, 3])
        yt = add(add(add(mul(T[1, 0], x), mul(T[1, 1], y)), mul(T[1, 2], z)), T[1, 3])
        d = np.sqrt(add(add(mul(x, x), mul(y, y)), mul(z, z))).astype(f32)
        vN = mul(f32(min_r), f32(min_r))
        b = add(f32(beam_c), mul(f32(beam_a), d))
        vL = mul(b, b)
        sJ = np.asarray(sJ, f32)
        zero = f32(0)
        B0 = add(add(mul(sJ[0], vL), mul(sJ[1], zero)), mul(sJ[2], zero))
        B1 = add(add(mul(sJ[0], zero), mul(sJ[1], vL)), mul(sJ[2], zero))
        B2 = add(add(mul(sJ[0], zero), mul(sJ[1], zero)), mul(sJ[2], vN) * np.ones_like(vL))
        hv = add(add(mul(B0, sJ[0]), mul(B1, sJ[1])), mul(B2, sJ[2]))
        hv = add(zero, hv)
    geo, sto = points_to_index(xt, yt, L, res, centre, start)
    key = np.where(acc, sto, -1).astype(np.int32)
    geo = np.where(acc, geo, -1).astype(np.int32)
    m1 = f32(-1)
    return (key, geo, np.where(acc, h, m1).astype(f32), np.where(acc, hv, m1).astype(f32),
            np.where(acc, xt, m1).astype(f32), np.where(acc, yt, m1).astype(f32))


def lowest_update(lowest, geo, h, var):
    """ORACLE DEFINITION of gpu.cu:432-438"""
    lowest = lowest.copy().reshape(-1)
    best = {}
    for i in range(len(geo)):
        g = int(geo[i])
        if g < 0:
            continue
        if g not in best or h[i] < h[best[g]]:
            best[g] = i
    for g, i in best.items():
        if h[i] <= lowest[g]:
            lowest[g] = f32(h[i]) + f32(f32(3) * f32(var[i]))
    return lowest


def fuse(elev, var, inten, cr, cg, cb, key, R, G, B, I, h, v):
    """gpu.cu:477-537, O(N) in index order; arrays are flat copies, returned updated"""
    elev = elev.copy(); var = var.copy(); inten = inten.copy(); cr = cr.copy(); cg = cg.copy(); cb = cb.copy()
    C = elev.size
    for i in range(len(key)):
        c = int(key[i])
        if c < 0 or c >= C or h[i] == f32(-1):
            continue
        hi, vi = f32(h[i]), f32(v[i])
        col = (R[i] != 0) and (G[i] != 0) and (B[i] != 0) and (I[i] != 0)
        take = False
        if elev[c] == f32(-10):
            elev[c] = hi; var[c] = vi; take = True
        else:
            if float(var[c]) < 0.0001:
                var[c] = f32(0.0001)
            with np.errstate(all="ignore"):
                mah = f32(abs(f32(hi - elev[c]))) / f32(np.sqrt(var[c]))
            if mah > 5:
                if elev[c] < hi:
                    elev[c] = hi; var[c] = vi; take = True
            else:
                ov, oe = var[c], elev[c]
                with np.errstate(all="ignore"):
                    elev[c] = f32(f32(f32(ov * hi) + f32(vi * oe)) / f32(ov + vi))
                    var[c] = f32(f32(vi * ov) / f32(vi + ov))
                take = True
        if take and col:
            inten[c] = I[i]; cr[c] = R[i]; cg[c] = G[i]; cb[c] = B[i]
    low = var.astype(np.float64) < 0.0001
    var[low] = f32(0.0001)
    return elev, var, inten, cr, cg, cb


def raytracing(elev, var, traver, lowest, L, start, sensorZ, thr=0.7):
    """gpu.cu:708-891 in scalar float32 Python; returns new elevation (flat)"""
    elev = elev.copy().reshape(-1)
    var = var.reshape(-1); traver = traver.reshape(-1); lowest = lowest.reshape(-1)
    out = elev.copy()
    robot = int(f32(L // 2 - 0.5)) if L % 2 == 0 else L // 2
    for i in range(L * L):
        if not (traver[i] < f32(thr) and elev[i] != f32(-10)):
            continue
        cx0, cy0 = i // L, i % L
        ox, oy = (cx0 + L - start[0]) % L, (cy0 + L - start[1]) % L
        inc0, inc1 = f32(ox - robot), f32(oy - robot)
        ix = 1 if inc0 > 0 else (0 if inc0 == 0 else -1)
        iy = 1 if inc1 > 0 else (0 if inc1 == 0 else -1)
        if ix == 0 or iy == 0:
            continue
        restrict = f32(elev[i])
        dis = f32(np.sqrt(f32(f32(inc0 * inc0) + f32(inc1 * inc1))))
        d0, d1 = f32(inc0 / dis), f32(inc1 / dis)
        if abs(inc0) > abs(inc1):
            t = 0.5 / float(inc0) * float(inc1)
        else:
            t = 0.5 / float(inc1) * float(inc0)
        threshold = f32(math.sqrt(0.25 + t * t))
        bx, by = f32(f32(ix) / f32(2)), f32(f32(iy) / f32(2))
        with np.errstate(all="ignore"):
            dnx, dny = f32(bx / d0), f32(by / d1)
        later = f32(0)
        cx, cy = ox, oy

        def probe(cx, cy, restrict):
            low = lowest[cx * L + cy]
            if low == f32(10):
                return restrict
            x1 = f32(cx - ox)
            x2 = f32(f32(cx) - f32(robot))
            h2 = f32(f32(sensorZ) - low)
            with np.errstate(all="ignore"):
                me = f32(low + f32(f32(h2 / x2) * x1))
            return me if me < restrict else restrict

        while 0 <= cx < L and 0 <= cy < L:
            if dnx > dny:
                if f32(dny - later) > threshold and cx != ox and cy != oy:
                    restrict = probe(cx, cy, restrict)
                cy += iy; by = f32(by + f32(iy)); later = dny; dny = f32(by / d1)
            elif dnx < dny:
                if f32(dnx - later) > threshold and cx != ox and cy != oy:
                    restrict = probe(cx, cy, restrict)
                cx += ix; bx = f32(bx + f32(ix)); later = dnx; dnx = f32(bx / d0)
            else:
                if f32(dnx - later) > threshold and cx != ox and cy != oy:
                    restrict = probe(cx, cy, restrict)
                cx += ix; cy += iy; bx = f32(bx + f32(ix)); by = f32(by + f32(iy)); later = dnx
                dnx = f32(bx / d0); dny = f32(by / d1)
        with np.errstate(all="ignore"):
            if f32(elev[i] - f32(f32(3) * f32(np.sqrt(var[i])))) > restrict:
                out[i] = f32(-10)
    return out
