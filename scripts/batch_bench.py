import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, gem_b200
frames = bench.gen_frames(16)
fobjs = [bench.laser_frame(fr) for fr in frames]
npts = [fr["xyzi"].shape[0] for fr in frames]
r = bench.run_multi_sensor(frames, fobjs, npts, 6571.2, nsens=8, K=int(os.environ.get("K", "30")))
print(r)
