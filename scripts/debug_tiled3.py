import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gem_b200
from gem_b200 import synth, tiled
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = torch.device("cuda", local)
L, res = 1024 * world, 0.05
scene = synth.make_scene()
F = 8
frames = [synth.hdl64_frame(1000 * rank + k, scene=scene) for k in range(F)]
ox, oy = tiled.sensor_offset(rank, world)
fobjs = []
for k, fr in enumerate(frames):
    T = fr["T"].copy(); T[0, 3] = ox + (k - F / 2.0); T[1, 3] = oy
    fobjs.append(gem_b200.make_frame(T, gem_b200.LaserSensorProcessor()))
npts = [fr["xyzi"].shape[0] for fr in frames]
xd = [torch.from_numpy(fr["xyzi"]).to(dev) for fr in frames]
rd = [torch.from_numpy(fr["rgba"]).to(dev) for fr in frames]
cap = ((max(npts) + 1023) // 1024) * 1024
tm = tiled.TiledElevationMap(L, res, max_points=max(1 << 21, world * cap), bucket_capacity=cap)
torch.cuda.synchronize()
for s in range(6):
    k = s % F
    tm.add(xd[k], rd[k], fobjs[k])
    tm.map.sync(); torch.cuda.synchronize()
    st = tm.map.stats()
    rec = tm.recv.cpu().numpy()
    g = rec[:, 0]
    v = g[g >= 0]
    u, c = np.unique(v, return_counts=True)
    i = np.argmax(c)
    snd = tm.send.cpu().numpy()[:, 0]
    print(f"rank {rank} step {s} n={npts[k]} counts={tm.counts.tolist()} stats={st} recv_valid={len(v)} top_gkey={u[i]} ({u[i]//L},{u[i]%L}) x{c[i]} "
          f"send_valid_per_bucket={[int((snd[o*cap:(o+1)*cap] >= 0).sum()) for o in range(world)]} recv_valid_per_src={[int((g[o*cap:(o+1)*cap] >= 0).sum()) for o in range(world)]}", flush=True)
dist.destroy_process_group()
