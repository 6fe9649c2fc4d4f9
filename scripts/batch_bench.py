"""BASELINE configs[4] shape on ONE GPU: 8 sensors x ~123 k points per launch into an 8192 x 8192 @ 0.05 m map through
gem_add_points_multi.  python scripts/batch_bench.py [steps] -- prints us/step, Gpoints/s and the algorithmic GB/s;
run under ncu for the DRAM bytes (profiles/r2_batch_*.txt)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import gem_b200
from bench import gen_frames, laser_frame, run_multi_sensor, load_peaks
K = int(sys.argv[1]) if len(sys.argv) > 1 else 200
frames = gen_frames(32)
fobjs = [laser_frame(fr) for fr in frames]
npts = [fr["xyzi"].shape[0] for fr in frames]
print(run_multi_sensor(frames, fobjs, npts, load_peaks()[0], K=K))
