"""Executable model of gem_tiled_step's schedule (gem_api.cu) and receive-buffer reuse (gem_route.cuh, "Receive buffers").

Not a test of the CUDA code: a small event simulation of what the code relies on.  Every rank issues one graph per call on
one stream (graph j starts when graph j-1 has completed); kernels of one graph run concurrently unless the graph has an
edge; a route kernel of step k writes buffer k % B on EVERY rank and raises flag k there when it ends; a bin kernel of
step k can only end once every peer's flag k is up, and reads its own buffer k % B; the folds of step k read the own
buffer k % B (intensities).  Safety: nobody writes buffer slot b of rank o for step k' while o still has a reader of
slot b for an earlier step k < k' (running or not yet started).  The model explores random interleavings.

depth 2 (default): graph j = {fold(j-1) || route(j) -> bin(j)}     -> three buffers suffice, two do not
depth 3          : graph j = {fold(j-2) || bin(j-1) || route(j)}   -> five buffers suffice, four do not
"""
import random

import pytest


def simulate(world, depth, nbuf, calls, rng):
    """returns None when the run was safe, else a description of the first violation"""
    # kernel states per rank: dict name -> [state] with state in {"todo", "run", "done"}
    graphs = []  # graphs[r][j] = {kernel: state}
    for _ in range(world):
        g = []
        for j in range(1, calls + depth + 1):      # the trailing calls drain the pipeline (gem_flush)
            ks = {}
            if j <= calls:
                ks[("route", j)] = "todo"
            if depth == 2:
                if j <= calls:
                    ks[("bin", j)] = "todo"
                if 1 <= j - 1 <= calls:
                    ks[("fold", j - 1)] = "todo"
            else:
                if 1 <= j - 1 <= calls:
                    ks[("bin", j - 1)] = "todo"
                if 1 <= j - 2 <= calls:
                    ks[("fold", j - 2)] = "todo"
            g.append(ks)
        graphs.append(g)
    cur = [0] * world                       # index of the rank's running graph
    flags = [[0] * world for _ in range(world)]   # flags[o][r] = last step rank r delivered to rank o
    written = [[0] * nbuf for _ in range(world)]  # step whose records are in buffer slot b of rank o
    writing = [[set() for _ in range(nbuf)] for _ in range(world)]

    def readers_pending(o, b, step):
        """EARLIER steps whose bin / fold on rank o still has to read (or is reading) slot b: their records would be lost"""
        out = []
        for ks in graphs[o][cur[o]:]:
            for (name, k), st in ks.items():
                if name in ("bin", "fold") and k % nbuf == b and k < step and st != "done":
                    out.append((name, k, st))
        return out

    while True:
        moves = []
        for r in range(world):
            if cur[r] >= len(graphs[r]):
                continue
            ks = graphs[r][cur[r]]
            if all(st == "done" for st in ks.values()):
                moves.append(("next", r, None))
                continue
            for key, st in ks.items():
                name, k = key
                if st == "todo":
                    if name == "bin" and depth == 2 and ks.get(("route", k)) != "done":
                        continue          # graph edge route -> bin
                    moves.append(("start", r, key))
                elif st == "run":
                    if name == "bin" and any(flags[r][p] < k for p in range(world)):
                        continue          # spins on the flags
                    moves.append(("end", r, key))
        if not moves:
            break
        what, r, key = rng.choice(moves)
        if what == "next":
            cur[r] += 1
            continue
        name, k = key
        ks = graphs[r][cur[r]]
        if what == "start":
            ks[key] = "run"
            if name == "route":
                for o in range(world):
                    bad = readers_pending(o, k % nbuf, k)
                    if bad:
                        return f"rank {r} route of step {k} writes slot {k % nbuf} of rank {o} while {bad} pending there"
                    writing[o][k % nbuf].add((r, k))
        else:
            ks[key] = "done"
            if name == "route":
                for o in range(world):
                    writing[o][k % nbuf].discard((r, k))
                    written[o][k % nbuf] = k
                    flags[o][r] = k
            elif name == "bin":
                if written[r][k % nbuf] != k or any(kk != k for (_, kk) in writing[r][k % nbuf]):
                    return f"rank {r} bin of step {k} read slot {k % nbuf} holding step {written[r][k % nbuf]}"
    assert all(c >= len(g) for c, g in zip(cur, graphs)), "deadlock in the model"
    return None


@pytest.mark.parametrize("depth,nbuf", [(2, 3), (2, 5), (3, 5)])
def test_shipped_buffer_counts_are_safe(depth, nbuf):
    rng = random.Random(1234 + 10 * depth + nbuf)
    for world in (1, 2, 3):
        for _ in range(300):
            assert simulate(world, depth, nbuf, calls=9, rng=rng) is None


@pytest.mark.parametrize("depth,nbuf", [(2, 2), (3, 4), (3, 3)])
def test_one_buffer_less_is_caught_by_the_model(depth, nbuf):
    rng = random.Random(99 + 10 * depth + nbuf)
    found = None
    for _ in range(3000):
        found = simulate(2, depth, nbuf, calls=9, rng=rng)
        if found:
            break
    assert found, "the model should find an interleaving that overwrites a buffer still to be read"
