"""Host-side logic of the tiled multi-GPU path on CPU: tile planning, ownership, and the
variable-size all-to-all (gloo, world_size 2) that must concatenate buckets in
(source rank, source order) -- the property that makes the tiled map equal the single-GPU map."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gem_b200 import tiled


def test_tile_plan_matches_baseline_configs():
    assert tiled.plan_tiles(1) == (1, 1)
    assert tiled.plan_tiles(2) == (1, 2)
    assert tiled.plan_tiles(4) == (2, 2)      # config 4: 4096^2 as 2x2 tiles of 2048^2
    assert tiled.plan_tiles(8) == (2, 4)      # config 5: 8192^2 as 2x4 tiles of 4096x2048
    assert tiled.tile_of_rank(3, 4, 4096) == (2048, 2048, 2048, 2048)
    assert tiled.tile_of_rank(5, 8, 8192) == (4096, 4096, 2048, 2048)
    # tiles partition the map
    for world, L in ((2, 2048), (4, 4096), (8, 8192), (4, 1001)):
        cover = np.zeros((L, L), np.int32)
        for r in range(world):
            r0, nr, c0, nc = tiled.tile_of_rank(r, world, L)
            cover[r0:r0 + nr, c0:c0 + nc] += 1
            gx, gy = np.meshgrid(np.arange(r0, r0 + nr, 97), np.arange(c0, c0 + nc, 89), indexing="ij")
            assert (tiled.owner_of(gx, gy, world, L) == r).all()
        assert (cover == 1).all()


def _worker(rank, world, port, tmp):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(100 + rank)
    L = 64
    n = 5000 + 777 * rank
    gx, gy = rng.integers(0, L, n), rng.integers(0, L, n)
    gkey = (gx * L + gy).astype(np.int32)
    gkey[rng.uniform(size=n) < 0.1] = -1                      # rejected / out of grid
    owner = np.where(gkey >= 0, tiled.owner_of(gkey // L, gkey % L, world, L), -1)
    rec = np.zeros((n, tiled.REC_WORDS), np.int32)
    rec[:, 0] = gkey
    rec[:, 1] = np.arange(n) + 1_000_000 * rank                 # payload: (source rank, source index)
    # stable bucket by owner == what k_route_count/scan/write do on the GPU
    order = np.concatenate([np.nonzero(owner == o)[0] for o in range(world)])
    counts = [int((owner == o).sum()) for o in range(world)]
    send = torch.from_numpy(rec[order])
    recv, out_splits = tiled.exchange(send, counts)
    recv = recv.numpy()
    # everything received belongs to this rank's tile
    r0, nr, c0, nc = tiled.tile_of_rank(rank, world, L)
    kx, ky = recv[:, 0] // L, recv[:, 0] % L
    assert ((kx >= r0) & (kx < r0 + nr) & (ky >= c0) & (ky < c0 + nc)).all()
    # grouped by source rank in rank order, source order preserved inside each group
    src = recv[:, 1] // 1_000_000
    assert (np.diff(src) >= 0).all()
    for s in range(world):
        idx = recv[src == s, 1]
        assert (np.diff(idx) > 0).all()
    assert sum(out_splits) == recv.shape[0]
    np.save(os.path.join(tmp, f"recv{rank}.npy"), recv)
    dist.barrier()
    dist.destroy_process_group()


def test_exchange_order_gloo_world2(tmp_path):
    port = 29600 + os.getpid() % 200
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    got = sum(np.load(tmp_path / f"recv{r}.npy").shape[0] for r in range(2))
    assert got > 9000


@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_halo_assembly_matches_slicing(world):
    """padded_from_borders / global_from_tiles (features and ray clean-up on tiled maps) against plain slicing"""
    import torch
    L = 16
    g = torch.arange(L * L, dtype=torch.float32).view(L, L)
    tiles = []
    for r in range(world):
        r0, nr, c0, nc = tiled.tile_of_rank(r, world, L)
        tiles.append(g[r0:r0 + nr, c0:c0 + nc].contiguous())
    assert torch.equal(tiled.global_from_tiles(tiles, world, L), g)
    borders = torch.stack([tiled.border_pack(t) for t in tiles])
    gp = torch.full((L + 4, L + 4), -10.0)
    gp[2:-2, 2:-2] = g
    for r in range(world):
        r0, nr, c0, nc = tiled.tile_of_rank(r, world, L)
        assert torch.equal(tiled.padded_from_borders(tiles[r], borders, r, world), gp[r0:r0 + nr + 4, c0:c0 + nc + 4])


def _halo_worker(rank, world, port):
    """the collective part of TiledElevationMap.compute_features / clean: all_gather of the border strips and of the
    lowest tiles, then local assembly -- with gloo on CPU tensors"""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    L = 32
    g = (torch.arange(L * L, dtype=torch.float32).view(L, L) * 0.5 - 100.0)
    r0, nr, c0, nc = tiled.tile_of_rank(rank, world, L)
    own = g[r0:r0 + nr, c0:c0 + nc].contiguous()
    mine = tiled.border_pack(own)
    flat = torch.empty(world * mine.numel(), dtype=torch.float32)
    dist.all_gather_into_tensor(flat, mine)
    padded = tiled.padded_from_borders(own, flat.view(world, mine.numel()), rank, world)
    gp = torch.full((L + 4, L + 4), -10.0)
    gp[2:-2, 2:-2] = g
    assert torch.equal(padded, gp[r0:r0 + nr + 4, c0:c0 + nc + 4])
    flat = torch.empty(world * nr * nc, dtype=torch.float32)
    dist.all_gather_into_tensor(flat, own.view(-1))
    assert torch.equal(tiled.global_from_tiles(list(flat.view(world, nr, nc)), world, L), g)
    dist.barrier()
    dist.destroy_process_group()


def test_halo_and_lowest_gather_gloo_world2():
    port = 29820 + os.getpid() % 150
    mp.spawn(_halo_worker, args=(2, port), nprocs=2, join=True)
