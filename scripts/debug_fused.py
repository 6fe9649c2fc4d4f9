import sys, os, time, ctypes as C
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import gem_b200
from gem_b200 import synth
from oracle_lib import OracleMap
sc = synth.make_scene()
frames = [synth.hdl64_frame(k, scene=sc) for k in range(6)]
fobjs = [gem_b200.make_frame(fr["T"], gem_b200.LaserSensorProcessor()) for fr in frames]
g = gem_b200.ElevationMap(1024, 0.05, compat_box_filter=False)
stream = g.torch_stream()
g.debug_phase_stamps(True)
o = OracleMap(1024, 0.05, compat_box_filter=False)
xd = [torch.from_numpy(fr["xyzi"]).cuda() for fr in frames]
rd = [torch.from_numpy(fr["rgba"]).cuda() for fr in frames]
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
NS = 30
seq = [s % 6 for s in range(NS)]
t0 = time.perf_counter()
e0.record(stream)
for k in seq:
    g.move(frames[k]["position"]); g.add(xd[k], rd[k], fobjs[k])
e1.record(stream)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print("enqueue ms", (t1 - t0) * 1e3, "sync ms", (t2 - t1) * 1e3, "event ms", e0.elapsed_time(e1), "per step us", e0.elapsed_time(e1) / NS * 1e3)
for k in seq:
    o.move(frames[k]["position"]); o.add(frames[k]["xyzi"], frames[k]["rgba"], fobjs[k])
for name in ("elevation", "variance", "intensity", "color_r", "lowest"):
    a, b = g.get_layer(name), o.get_layer(name)
    print(name, "mismatch", int((a != b).sum()), "valid", int((b != -10).sum()))
print(g.stats())
g.move(frames[1]["position"]); g.add(xd[1], rd[1], fobjs[1]); g.sync()
st = g.debug_phase_stamps(True)
print("phase stamps (us from start):", [round((t - st[0]) / 1e3, 2) for t in st[:11]])
print("last block end us", (st[9] - st[0]) / 1e3)
pr_names = ["transform", "sync1", "alloc", "sync2", "scatter", "sync3", "fold_large", "fold_small"]
print({n: round((st[i + 1] - st[i]) / 1e3, 2) for i, n in enumerate(pr_names)})
g.profile_enable(True)
for k in seq:
    g.move(frames[k]["position"]); g.add(xd[k], rd[k], fobjs[k])
pr = g.profile_read()
print({k: round(v / max(1, pr["count"][k]) * 1e3, 2) for k, v in pr["ms"].items() if pr["count"][k]})
