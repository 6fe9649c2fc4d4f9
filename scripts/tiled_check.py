"""torchrun worker: N ranks build a tiled map over NCCL and compare it with the untiled map that
rank 0 computes alone on the same clouds (bit for bit)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gem_b200  # noqa: E402
from gem_b200 import synth, tiled  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
L, res, steps = 512 * world, 0.1, 3
scene = synth.make_scene()
dev = torch.device("cuda", local)


def cloud(r, s):
    fr = synth.hdl64_frame(10 * r + s, scene=scene)
    ox, oy = tiled.sensor_offset(r, world)
    fr["T"] = fr["T"].copy()
    fr["T"][:2, 3] = (ox * 0.5 + s, oy * 0.5)
    return fr, gem_b200.make_frame(fr["T"], gem_b200.LaserSensorProcessor())


mode = os.environ.get("TILED_MODE", "peer")   # peer | padded | packed
tm = tiled.TiledElevationMap(L, res, max_points=1 << 20, bucket_capacity=0 if mode == "packed" else (1 << 17) + 4096,
                             peer=(mode == "peer"))
for s in range(steps):
    fr, f = cloud(rank, s)
    tm.add(torch.from_numpy(fr["xyzi"]).to(dev), torch.from_numpy(fr["rgba"]).to(dev), f)
pos = np.array([0.0, 0.0, 1.8], np.float32)
tm.map.move(pos)
tm.compute_features()      # halo all-gather + 5x5 PCA on the padded tile
tm.clean()                 # replicated lowest + ray clean-up of the own tile
tm.map.sync()
torch.cuda.synchronize()
ok = True
NAMES = ("elevation", "variance", "intensity", "color_r", "traver", "rough", "slope", "lowest")
if rank == 0:
    single = gem_b200.ElevationMap(L, res, compat_box_filter=False)
    single.move(pos)
    for s in range(steps):          # per step ONE multi-sensor frame: the ranks' clouds in rank order
        cl = [cloud(r, s) for r in range(world)]
        xa = torch.cat([torch.from_numpy(c[0]["xyzi"]) for c in cl]).to(dev)
        ca = torch.cat([torch.from_numpy(c[0]["rgba"]) for c in cl]).to(dev)
        offs = np.concatenate([[0], np.cumsum([c[0]["xyzi"].shape[0] for c in cl])])
        single.add_multi(xa, ca, offs, [c[1] for c in cl])
        single.sync()
    single.compute_features()
    single.raytracing()
    full = {n: single.get_layer(n) for n in NAMES}
for name in NAMES:
    mine = torch.from_numpy(tm.get_layer(name).astype(np.float32).copy()).to(dev)
    gathered = [torch.empty_like(mine) for _ in range(world)] if rank == 0 else None
    dist.gather(mine, gathered, dst=0)
    if rank == 0:
        for r in range(world):
            r0, nr, c0, nc = tiled.tile_of_rank(r, world, L)
            a = gathered[r].cpu().numpy()
            b = full[name][r0:r0 + nr, c0:c0 + nc].astype(np.float32)
            if not np.array_equal(a.view(np.uint32), b.view(np.uint32)):
                ok = False
                print("MISMATCH", name, r, int((a != b).sum()))
if rank == 0:
    valid = int((full["elevation"] != -10).sum())
    print("valid cells", valid)
    print("TILED_CHECK_OK" if ok and valid > 10000 else "TILED_CHECK_FAILED")
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
