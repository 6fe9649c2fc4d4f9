"""Tiled map on the GPU.

test_two_tiles_on_one_gpu_equal_single_map runs in the single-GPU tier: two tile handles on one
device, the all-to-all replaced by slicing, must reproduce the untiled map bit for bit.
test_nccl_two_ranks needs >= 2 GPUs (run with `gpurun --gpus 2`): real NCCL all-to-all."""
import os
import subprocess
import sys

import numpy as np
import pytest

import gem_b200
from gem_b200 import synth, tiled

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LAYERS = ["elevation", "variance", "intensity", "color_r", "color_g", "color_b"]


def test_two_tiles_on_one_gpu_equal_single_map():
    import torch
    L, res, world = 512, 0.1, 2
    scene = synth.make_scene()
    frs = [synth.hdl64_frame(k, scene=scene) for k in range(2)]
    for k, fr in enumerate(frs):        # two sensors 20 m apart
        fr["T"] = fr["T"].copy()
        fr["T"][:2, 3] = (-10.0 + 20.0 * k, 3.0 * k)
    fobj = [gem_b200.make_frame(fr["T"], gem_b200.LaserSensorProcessor()) for fr in frs]
    single = gem_b200.ElevationMap(L, res, compat_box_filter=False)
    for fr, f in zip(frs, fobj):
        single.add(fr["xyzi"], fr["rgba"], f)
    tiles = [gem_b200.ElevationMap(L, res, compat_box_filter=False, tile=tiled.tile_of_rank(r, world, L)) for r in range(world)]
    tr, tc = tiled.plan_tiles(world)
    dev = torch.device("cuda", 0)
    sends, counts = [], []
    for r in range(world):              # "rank r" routes its own sensor's cloud
        x = torch.from_numpy(frs[r]["xyzi"]).to(dev)
        c = torch.from_numpy(frs[r]["rgba"]).to(dev)
        send = torch.zeros((x.shape[0], tiled.REC_WORDS), dtype=torch.int32, device=dev)
        cnt = torch.zeros(world, dtype=torch.int32, device=dev)
        tiles[r].route_points(x, c, fobj[r], tr, tc, send, cnt)
        tiles[r].sync()
        sends.append(send)
        counts.append(cnt.cpu().tolist())
    for dst in range(world):            # all-to-all by slicing: (source rank, source order)
        parts = []
        for src in range(world):
            off = sum(counts[src][:dst])
            parts.append(sends[src][off:off + counts[src][dst]])
        recv = torch.cat(parts).contiguous()
        tiles[dst].fuse_records(recv, recv.shape[0])
        tiles[dst].sync()
    total = 0
    for name in LAYERS:
        full = single.get_layer(name)
        for r in range(world):
            r0, nr, c0, nc = tiled.tile_of_rank(r, world, L)
            a, b = tiles[r].get_layer(name), full[r0:r0 + nr, c0:c0 + nc]
            same = (a == b) if a.dtype.kind != "f" else (a.view(np.uint32) == b.view(np.uint32))
            assert same.all(), (name, r)
        total += int((full != (-10 if name in ("elevation",) else 0)).sum()) if name == "elevation" else 0
    assert total > 20000
    assert sum(sum(c) for c in counts) == sum(t.stats()["points_binned"] for t in tiles)


def test_nccl_two_ranks():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29655", os.path.join(ROOT, "scripts", "tiled_check.py")]
    for mode in ("peer", "padded", "packed"):   # NVLink peer stores / padded NCCL all-to-all / packed NCCL
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=dict(os.environ, TILED_MODE=mode))
        print(r.stdout[-3000:], r.stderr[-3000:])
        assert r.returncode == 0 and "TILED_CHECK_OK" in r.stdout
