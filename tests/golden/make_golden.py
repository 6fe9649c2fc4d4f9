#!/usr/bin/env python
"""Generate the committed golden fixture tests/golden/gem_golden_v1.npz.

Inputs: a seeded 3-frame HDL-64E-shaped stream (subsampled to keep the file small), reference
demo axes (so the hard-coded box filter of gpu_process.cu:393 keeps points), 96x96 @ 0.2 m map.
Outputs per frame, produced by THE REFERENCE ITSELF when run where oracle/_ref exists and a GPU
is present (`gpurun -- python tests/golden/make_golden.py`): Move outputs, Process_points
outputs, and the layers after Fuse / Map_feature / Raytracing, from
  ref_nofma : /root/reference/.../gpu_process.cu compiled unmodified with -fmad=false
  ref_fma   : the same file with the reference's own flags (FMA contraction on)
plus the CPU oracle's outputs on the same inputs.  The file is written to gpurun_out/ on the GPU
box and copied to tests/golden/ by hand (the GPU box is not the repository).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import gem_b200  # noqa: E402
from gem_b200 import synth  # noqa: E402
from oracle_lib import OracleMap  # noqa: E402

L, RES, NFRAMES, STRIDE = 96, 0.2, 3, 11


def inputs():
    scene = synth.make_scene()
    out = []
    for k in range(NFRAMES):
        fr = synth.hdl64_frame(k, scene=scene, compat_axes=True, speed=8.0)
        fr["xyzi"] = np.ascontiguousarray(fr["xyzi"][k::STRIDE])
        fr["rgba"] = np.ascontiguousarray(fr["rgba"][k::STRIDE])
        fr["rgba"][::13, 1] = 0          # exercise the "any channel zero -> keep old colour" rule
        out.append(fr)
    return out


def run(m, frames, is_ref, lowest_from=None):
    """drive one implementation; returns dict of arrays"""
    res = {}
    for k, fr in enumerate(frames):
        f = gem_b200.make_frame(fr["T"], gem_b200.LaserSensorProcessor())
        centre, start, shift = m.move(fr["position"])
        x, y, z = (fr["xyzi"][:, j] for j in range(3))
        key, var, xt, yt, zt = m.process_points(x, y, z, f)
        R, G, B = (fr["rgba"][:, j].astype(np.int32) for j in range(3))
        m.fuse_points(key, R, G, B, fr["xyzi"][:, 3], zt, var)
        if is_ref and lowest_from is not None:
            # the reference's `lowest` update is a data race; use the oracle's definition so the
            # ray step is comparable (documented in tests/test_reference_pin.py)
            m.set_layer("lowest", lowest_from[k])
        feat = m.map_feature()
        m.raytracing()
        res[f"f{k}_centre"], res[f"f{k}_start"], res[f"f{k}_shift"] = centre, start, shift
        res[f"f{k}_key"], res[f"f{k}_var"], res[f"f{k}_xt"], res[f"f{k}_yt"], res[f"f{k}_zt"] = key, var, xt, yt, zt
        for name in ("elevation", "variance", "intensity", "color_r", "color_g", "color_b", "rough", "slope", "traver"):
            res[f"f{k}_feat_{name}"] = feat[name]
        res[f"f{k}_elev_after_ray"] = m.get_layer("elevation").reshape(-1)
    return res


def main():
    frames = inputs()
    data = {"L": L, "res": RES, "nframes": NFRAMES}
    for k, fr in enumerate(frames):
        data[f"in{k}_xyzi"], data[f"in{k}_rgba"], data[f"in{k}_T"], data[f"in{k}_pos"] = fr["xyzi"], fr["rgba"], fr["T"], fr["position"]
    # oracle (also records its lowest layer after process_points for the reference's ray step)
    o = OracleMap(L, RES, compat_box_filter=True)
    lows = []

    class Spy:
        def __getattr__(self, n):
            return getattr(o, n)

        def fuse_points(self, *a):
            lows.append(o.get_layer("lowest").copy())
            return o.fuse_points(*a)
    ores = run(Spy(), frames, False)
    data.update({"oracle_" + k: v for k, v in ores.items()})
    import ref_lib
    have_gpu = False
    try:
        import torch
        have_gpu = torch.cuda.is_available()
    except Exception:
        pass
    if have_gpu and ref_lib.available(True):
        for tag, nofma in (("ref_nofma", True), ("ref_fma", False)):
            r = ref_lib.RefMap(L, RES, nofma=nofma)
            rres = run(r, frames, True, lowest_from=lows)
            data.update({tag + "_" + k: v for k, v in rres.items()})
        data["generated_by"] = "reference gpu_process.cu on " + torch.cuda.get_device_name(0)
        outdir = os.path.join(ROOT, "gpurun_out")
    else:
        data["generated_by"] = "oracle only (no GPU / oracle/_ref)"
        outdir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(outdir, exist_ok=True)
    path = os.path.join(outdir, "gem_golden_v1.npz")
    np.savez_compressed(path, **data)
    print("wrote", path, os.path.getsize(path), "bytes;", data["generated_by"])


if __name__ == "__main__":
    main()
