import sys, os, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gem_b200
from gem_b200 import synth, tiled
L, res, world = 2048, 0.05, 2
fr = synth.hdl64_frame(0)
f = gem_b200.make_frame(fr["T"], gem_b200.LaserSensorProcessor())
dev = torch.device("cuda", 0)
x = torch.from_numpy(fr["xyzi"]).to(dev); c = torch.from_numpy(fr["rgba"]).to(dev)
cap = ((x.shape[0] + 1023) // 1024) * 1024
t = gem_b200.ElevationMap(L, res, compat_box_filter=False, tile=tiled.tile_of_rank(0, world, L), max_points=1 << 21)
send = torch.zeros((world * cap, 5), dtype=torch.int32, device=dev)
cnt = torch.zeros(world, dtype=torch.int32, device=dev)
for stride in (0, cap):
    for rep in range(3):
        t.profile_read(reset=True); t.profile_enable(True)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        t.route_points(x, c, f, 1, 2, send, cnt, stride)
        t.sync(); t1 = time.perf_counter()
        n = world * cap if stride else int(cnt.sum().item())
        t.fuse_records(send, n)
        t.sync(); t2 = time.perf_counter()
        pr = t.profile_read(reset=True); t.profile_enable(False)
        print("stride", stride, "route ms", (t1 - t0) * 1e3, "fuse ms", (t2 - t1) * 1e3, "n", n, cnt.tolist(),
              {k: round(v, 3) for k, v in pr["ms"].items() if v}, t.stats())
