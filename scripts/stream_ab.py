"""A/B of the stream-mode frame time (c2 workload) under an environment switch read at gem_create.
usage: python scripts/stream_ab.py GEM_B200_PDL_FRONT 0 1"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gem_b200  # noqa: E402
from gem_b200 import synth  # noqa: E402

var, values = sys.argv[1], sys.argv[2:]
L, res, NF, K = 1024, 0.05, 64, 600
scene = synth.make_scene()
frames = [synth.hdl64_frame(k, scene=scene) for k in range(NF)]
fobjs = [gem_b200.make_frame(fr["T"], gem_b200.LaserSensorProcessor()) for fr in frames]
xd = [torch.from_numpy(fr["xyzi"]).cuda() for fr in frames]
rd = [torch.from_numpy(fr["rgba"]).cuda() for fr in frames]
pos = [(C.c_float * 3)(*[float(v) for v in fr["position"]]) for fr in frames]
npts = sum(int(x.shape[0]) for x in xd) / NF
for rep in range(2):
    for v in values:
        os.environ[var] = v
        m = gem_b200.ElevationMap(L, res, compat_box_filter=False)
        st = m.torch_stream()

        def step(i):
            k = i % NF
            m.move_fast(pos[k])
            m.add_stream_fast(C.c_void_p(xd[k].data_ptr()), C.c_void_p(rd[k].data_ptr()), int(xd[k].shape[0]), C.byref(fobjs[k]))
        for i in range(40):
            step(i)
        m.sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(st):
            e0.record()
            for i in range(K):
                step(40 + i)
            e1.record()
        m.sync()
        us = e0.elapsed_time(e1) / K * 1e3
        print(f"{var}={v} rep{rep}: {us:.2f} us/frame  {npts / us:.0f} Mpoints/s", flush=True)
        m.close()
