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
        torch.cuda.synchronize()
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
        torch.cuda.synchronize()
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


def _add_round(single, tiles, world, L, frs, fobj):
    """one step: every 'rank' routes its sensor's cloud, buckets exchanged by slicing, owners fold"""
    import torch
    dev = torch.device("cuda", 0)
    tr, tc = tiled.plan_tiles(world)
    # a tiled step is ONE frame made of every rank's cloud (rank order): the per-frame `lowest` layer is the
    # minimum over all of them, which is what gem_add_points_multi computes on one GPU
    xa = torch.cat([torch.from_numpy(fr["xyzi"]) for fr in frs]).to(dev)
    ca = torch.cat([torch.from_numpy(fr["rgba"]) for fr in frs]).to(dev)
    offs = np.concatenate([[0], np.cumsum([fr["xyzi"].shape[0] for fr in frs])])
    single.add_multi(xa, ca, offs, fobj)
    single.sync()
    sends, counts = [], []
    for r in range(world):
        x = torch.from_numpy(frs[r]["xyzi"]).to(dev)
        c = torch.from_numpy(frs[r]["rgba"]).to(dev)
        send = torch.zeros((x.shape[0], tiled.REC_WORDS), dtype=torch.int32, device=dev)
        cnt = torch.zeros(world, dtype=torch.int32, device=dev)
        torch.cuda.synchronize()
        tiles[r].route_points(x, c, fobj[r], tr, tc, send, cnt)
        tiles[r].sync()
        sends.append(send)
        counts.append(cnt.cpu().tolist())
    for dst in range(world):
        parts = [sends[src][sum(counts[src][:dst]):sum(counts[src][:dst]) + counts[src][dst]] for src in range(world)]
        recv = torch.cat(parts).contiguous()
        torch.cuda.synchronize()
        tiles[dst].fuse_records(recv, recv.shape[0])
        tiles[dst].sync()


def _tile_layer(t, name, rows, cols):
    import torch
    out = torch.empty((rows, cols), dtype=torch.int32 if name.startswith("color") else torch.float32, device="cuda:0")
    t.get_layer_device(name, out)
    t.sync()
    return out


def _assert_tiles_equal(single_arrays, tiles, world, L, names, what):
    for name in names:
        full = single_arrays[name]
        for r in range(world):
            r0, nr, c0, nc = tiled.tile_of_rank(r, world, L)
            a = _tile_layer(tiles[r], name, nr, nc).cpu().numpy()
            b = full[r0:r0 + nr, c0:c0 + nc]
            same = (a.view(np.uint32) == np.ascontiguousarray(b).view(np.uint32)) | ((a != a) & (b != b))
            assert same.all(), f"{what}: {name} tile {r}: {np.count_nonzero(~same)} cells differ"


@pytest.mark.parametrize("world", [2, 4])
def test_tiled_features_and_cleanup_equal_single_map(world):
    """Map_feature (2-cell halo from the neighbouring tiles) and Raytracing (replicated lowest layer) on tile
    handles reproduce the untiled map cell for cell; world=4 (2x2) exercises the corner halos."""
    import torch
    L, res = 512, 0.1
    scene = synth.make_scene()
    single = gem_b200.ElevationMap(L, res, compat_box_filter=False)
    tiles = [gem_b200.ElevationMap(L, res, compat_box_filter=False, tile=tiled.tile_of_rank(r, world, L)) for r in range(world)]
    pos = np.array([0.0, 0.0, 1.8], np.float32)
    for m in [single] + tiles:
        m.move(pos)
    removed_total = 0
    for step in range(3):
        frs = [synth.hdl64_frame(step * world + k, scene=scene) for k in range(world)]
        for k, fr in enumerate(frs):            # sensors spread around the map centre (where the tile seams meet)
            fr["T"] = fr["T"].copy()
            fr["T"][:2, 3] = (-6.0 + 12.0 * (k % 2), -5.0 + 10.0 * (k // 2))
        fobj = [gem_b200.make_frame(fr["T"], gem_b200.LaserSensorProcessor()) for fr in frs]
        _add_round(single, tiles, world, L, frs, fobj)
        state = {n: single.get_layer(n) for n in LAYERS + ["lowest"]}
        _assert_tiles_equal(state, tiles, world, L, LAYERS + ["lowest"], f"step {step} add")

        # ---- features
        single.compute_features()
        tile_elev = [_tile_layer(tiles[r], "elevation", *[tiled.tile_of_rank(r, world, L)[i] for i in (1, 3)]) for r in range(world)]
        borders = torch.stack([tiled.border_pack(e) for e in tile_elev])
        for r in range(world):
            padded = tiled.padded_from_borders(tile_elev[r], borders, r, world).contiguous()
            assert np.array_equal(padded[2:-2, 2:-2].cpu().numpy(), tile_elev[r].cpu().numpy())
            tiles[r].compute_features_tiled(padded)
            tiles[r].sync()
        feat = {"rough": single.get_layer("rough"), "slope": single.get_layer("slope"), "traver": single.get_layer("traver")}
        _assert_tiles_equal(feat, tiles, world, L, ["rough", "slope", "traver"], f"step {step} features")
        # the halo must have mattered: cells on the seams have valid neighbours on the other side
        assert (feat["traver"] != -10).sum() > 5000

        # ---- a "dynamic obstacle" that left: raise a block of cells so later rays pass under their tops
        if step == 1:
            e = single.get_layer("elevation").copy()
            blk = (slice(L // 2 - 30, L // 2 + 30), slice(L // 2 - 30, L // 2 + 30))
            bump = np.where(e[blk] != -10, e[blk] + np.float32(1.5), e[blk]).astype(np.float32)
            e[blk] = bump
            single.set_layer("elevation", e)
            for r in range(world):
                r0, nr, c0, nc = tiled.tile_of_rank(r, world, L)
                tiles[r].set_layer("elevation", e[r0:r0 + nr, c0:c0 + nc])

        # ---- ray clean-up
        before = single.get_layer("elevation")
        single.raytracing()
        tile_low = [_tile_layer(tiles[r], "lowest", *[tiled.tile_of_rank(r, world, L)[i] for i in (1, 3)]) for r in range(world)]
        glob = tiled.global_from_tiles(tile_low, world, L)
        assert np.array_equal(glob.cpu().numpy().view(np.uint32), state["lowest"].view(np.uint32))
        for r in range(world):
            tiles[r].raytracing_tiled(glob)
            tiles[r].sync()
        after = {n: single.get_layer(n) for n in LAYERS + ["lowest"]}
        removed_total += int(((before != -10) & (after["elevation"] == -10)).sum())
        _assert_tiles_equal(after, tiles, world, L, LAYERS + ["lowest"], f"step {step} clean")
    print("cells removed by the ray clean-up:", removed_total)
    assert removed_total > 100
