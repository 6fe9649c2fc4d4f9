import sys, os, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gem_b200
from gem_b200 import synth
L, res = 1024, 0.05
scene = synth.make_scene()
frames = [synth.hdl64_frame(k, scene=scene) for k in range(12)]
fobjs = [gem_b200.make_frame(fr["T"], gem_b200.LaserSensorProcessor()) for fr in frames]
m = gem_b200.ElevationMap(L, res, compat_box_filter=False)
xd = [torch.from_numpy(fr["xyzi"]).cuda() for fr in frames]
rd = [torch.from_numpy(fr["rgba"]).cuda() for fr in frames]
ex = {n: np.empty((L, L), np.float32, order="F") for n in gem_b200._lib.EXPORT_LAYERS}
exp = {n: torch.empty((L, L), dtype=torch.float32).pin_memory() for n in gem_b200._lib.EXPORT_LAYERS}
exp_np = {n: exp[n].numpy().T for n in exp}   # F-ordered views of pinned memory
torch.cuda.synchronize()
def frame(k, out):
    m.move(frames[k]["position"]); m.add(xd[k], rd[k], fobjs[k]); m.var_update(0.0)
    m.compute_features(); m.export_layers(out); m.raytracing()
for k in range(6): frame(k, ex)
for name, out in (("pageable", ex), ("pinned", exp_np)):
    m.profile_read(reset=True); m.profile_enable(True)
    t0 = time.perf_counter()
    for k in range(6, 12): frame(k, out)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 6 * 1e3
    pr = m.profile_read(reset=True); m.profile_enable(False)
    print(name, "frame ms", round(dt, 3), {k: round(v / 6 * 1e3, 1) for k, v in pr["ms"].items() if v}, "(us per frame)")
valid = int((m.get_layer("elevation") != -10).sum()); print("valid cells", valid)
