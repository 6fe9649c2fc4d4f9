"""Timing of the SURVEY 8f rows next to the hot path (c2 map: L=1024, res 0.05, HDL-64 frames):
f1 layer export / orthomosaic / visual cloud, f2 colourisation, f3 prevMap_ snapshot + scroll-out harvest.
Device time = CUDA events around each kernel (gem_profile_*); wall = the host-synchronous C-ABI call incl. D2H.
Writes gpurun_out/next_rows.json; summarised in profiles/r1_next_rows.md."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gem_b200  # noqa: E402
from gem_b200 import synth  # noqa: E402

L, res, NF = 1024, 0.05, 16
scene = synth.make_scene()
frames = [synth.hdl64_frame(k, scene=scene) for k in range(NF)]
fobjs = [gem_b200.make_frame(fr["T"], gem_b200.LaserSensorProcessor()) for fr in frames]
m = gem_b200.ElevationMap(L, res, compat_box_filter=False, grid_resolution=0.05)
xd = [torch.from_numpy(fr["xyzi"]).cuda() for fr in frames]
rd = [torch.from_numpy(fr["rgba"]).cuda() for fr in frames]
C = L * L
peaks = {}
try:
    peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
except Exception:
    pass
hbm = float(peaks.get("hbm_gbs", 6571.2))
rows = {}


def timed(name, fn, bytes_algo, reps):
    """fn(k) is host-synchronous; device ms from the per-kernel profile"""
    m.sync()
    m.profile_read(reset=True)
    m.profile_enable(True)
    t0 = time.perf_counter()
    for k in range(reps):
        fn(k)
    m.sync()
    wall = (time.perf_counter() - t0) / reps * 1e6
    pr = m.profile_read(reset=True)
    m.profile_enable(False)
    dev = sum(pr["ms"].values()) / reps * 1e3
    rows[name] = {"device_us": round(dev, 2), "wall_us": round(wall, 1), "launches": pr["launches"] // reps,
                  "algorithmic_bytes": int(bytes_algo), "achieved_gbs": round(bytes_algo / (dev * 1e-6) / 1e9, 1) if dev else None,
                  "frac_of_hbm_peak": round(bytes_algo / (dev * 1e-6) / 1e9 / hbm, 4) if dev else None}


# populate: drive 10 frames with the full per-frame sequence
cur = shift = None
for k in range(10):
    cur, _, shift = m.move(frames[k]["position"])
    m.add(xd[k], rd[k], fobjs[k])
    m.compute_features()
    m.snapshot_shown()
    m.raytracing()
m.compute_features()
shown = int(m.export_visual_points(capacity=0)[2])
ex = {n: np.empty((L, L), np.float32, order="F") for n in gem_b200._lib.EXPORT_LAYERS}

timed("f1 export 9 layers (k_export_colmajor + 37.7 MB D2H)", lambda k: m.export_layers(ex), 20 * C + 36 * C, 5)
timed("f1 orthomosaic (k_orthomosaic + 3.1 MB D2H)", lambda k: m.export_orthomosaic(), 12 * C + 8 * shown + 3 * C, 10)
timed("f1 visual cloud (count+scan+write + D2H)", lambda k: m.export_visual_points(), 2 * 12 * C + (8 + 15) * shown, 10)
timed("f3 snapshot prevMap_ (k_snapshot_shown, D2D)", lambda k: m.snapshot_shown(), 40 * C, 10)

# harvest: move one frame ahead (1 m = 20 cells), harvest against the snapshot
harv = []


m.sync()
for k in range(5):
    c, _, s = m.move(frames[10 + k]["position"])
    m.add(xd[10 + k], rd[10 + k], fobjs[10 + k])
    m.compute_features()
    m.snapshot_shown()
    cn, _, sn = m.move(frames[11 + k]["position"])
    m.sync()
    m.profile_read(reset=True)
    m.profile_enable(True)
    t0 = time.perf_counter()
    rec, n = m.harvest_scrolled_out(cn, sn)
    wall = (time.perf_counter() - t0) * 1e6
    pr = m.profile_read(reset=True)
    m.profile_enable(False)
    harv.append((n, sum(pr["ms"].values()) * 1e3, wall))
n_h = float(np.mean([h[0] for h in harv]))
dev = float(np.median([h[1] for h in harv]))
rows["f3 harvest (count+scan+write + D2H of harvested records)"] = {
    "device_us": round(dev, 2), "wall_us": round(float(np.median([h[2] for h in harv])), 1), "launches": 3,
    "algorithmic_bytes": int(2 * 4 * C + 52 * n_h), "harvested_cells_per_step": n_h,
    "achieved_gbs": round((2 * 4 * C + 52 * n_h) / (dev * 1e-6) / 1e9, 1), "frac_of_hbm_peak": round((2 * 4 * C + 52 * n_h) / (dev * 1e-6) / 1e9 / hbm, 4)}

# f2 colourisation: KITTI-sized image, lidar cloud
H, W = 376, 1241
img = torch.randint(0, 256, (H, W, 3), dtype=torch.uint8, device="cuda")
Tc = np.array([[721.5, 0, 609.5, 0], [0, 721.5, 172.8, 0], [0, 0, 1, 0]], np.float64)
Tl = np.array([[0, -1, 0, 0], [0, 0, -1, -0.08], [1, 0, 0, -0.27], [0, 0, 0, 1]], np.float64)
xs = xd[0].clone()
out = torch.empty((xs.shape[0], 4), dtype=torch.uint8, device="cuda")
n = int(xs.shape[0])
timed("f2 colourise %d points from a %dx%d image" % (n, W, H), lambda k: (m.colourise(xs, Tc, Tl, img, out), m.sync()), n * (16 + 3 + 4 + 4), 20)

res_json = {"config": {"workload": "c2 map L=1024 res=0.05, HDL-64 frames", "shown_cells": shown, "hbm_peak_gbs": hbm}, "rows": rows}
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res_json, open("gpurun_out/next_rows.json", "w"), indent=1)
print(json.dumps(res_json, indent=1))
