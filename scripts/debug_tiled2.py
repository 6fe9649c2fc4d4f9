import os, sys, time
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gem_b200
from gem_b200 import synth, tiled
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = torch.device("cuda", local)
L, res = 1024 * world, 0.05
fr = synth.hdl64_frame(rank)
f = gem_b200.make_frame(fr["T"], gem_b200.LaserSensorProcessor())
x = torch.from_numpy(fr["xyzi"]).to(dev); c = torch.from_numpy(fr["rgba"]).to(dev)
cap = ((x.shape[0] + 4095) // 1024) * 1024
tm = tiled.TiledElevationMap(L, res, max_points=1 << 21, bucket_capacity=cap)
def t(fn, n=20):
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e6
def route():
    with torch.cuda.stream(tm.stream):
        tm.map.route_points(x, c, f, tm.tiles_r, tm.tiles_c, tm.send, tm.counts, tm.cap)
def a2a():
    with torch.cuda.stream(tm.stream):
        dist.all_to_all_single(tm.recv, tm.send)
def a2a_default():
    dist.all_to_all_single(tm.recv, tm.send)
def fuse():
    with torch.cuda.stream(tm.stream):
        tm.map.fuse_records(tm.recv, tm.recv.shape[0])
def full():
    tm.add(x, c, f)
for name, fn in (("route", route), ("a2a", a2a), ("a2a_default_stream", a2a_default), ("fuse", fuse), ("full", full), ("full", full)):
    us = t(fn)
    if rank == 0: print(name, round(us, 1), "us")
dist.destroy_process_group()
