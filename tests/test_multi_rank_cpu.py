"""N > 1 host-side logic on CPU: world_size-2 gloo run of the slab planning used by decomposed runs.

Each rank builds the particles-per-plane histogram of ITS half of a synthetic dam-break column distribution, the
histograms are all-reduced over gloo, and both ranks must derive identical, balanced cuts with the host-layer planner
(include/sphinxsys_ck/slab_decomposition.h::planSlabCuts, no GPU needed).
"""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    from sphinxsys_b200 import host
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    planes = 341
    rng = np.random.default_rng(1234)  # same stream on every rank
    x_plane = np.clip((rng.uniform(0.0, 2.0, size=200_000) / 0.01625).astype(np.int64) + 2, 0, planes - 1)
    mine = x_plane[rank::world]
    hist = torch.from_numpy(np.bincount(mine, minlength=planes).astype(np.int64))
    dist.all_reduce(hist)
    cuts = host.plan_slab_cuts(hist.numpy().astype(np.uint64), 4)
    gathered = [None] * world
    dist.all_gather_object(gathered, cuts.tolist())
    q.put((rank, gathered, hist.numpy().tolist()))
    dist.destroy_process_group()


def test_slab_cuts_identical_and_balanced_across_ranks():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, gathered, hist in results:
        assert gathered[0] == gathered[1], "ranks disagree on the cuts"
        cuts = gathered[0]
        assert cuts[0] == 0 and cuts[-1] == 341 and all(b > a for a, b in zip(cuts, cuts[1:]))
        counts = [sum(hist[a:b]) for a, b in zip(cuts, cuts[1:])]
        assert sum(counts) == 200_000
        assert max(counts) - min(counts) <= 2 * max(hist), "each slab within one plane of the ideal share"


@pytest.mark.parametrize("nranks", [1, 2, 3, 8])
def test_plan_slab_cuts_edge_cases(nranks):
    from sphinxsys_b200 import host
    # all particles in one plane: every rank still gets at least one plane
    h = np.zeros(16, dtype=np.uint64)
    h[5] = 1000
    cuts = host.plan_slab_cuts(h, nranks)
    assert cuts[0] == 0 and cuts[-1] == 16 and np.all(np.diff(cuts) >= 1)
    # uniform: equal plane counts
    cuts = host.plan_slab_cuts(np.full(16, 10, dtype=np.uint64), nranks)
    if 16 % nranks == 0:
        assert np.all(np.diff(cuts) == 16 // nranks)
    with pytest.raises(Exception):
        host.plan_slab_cuts(np.ones(2, dtype=np.uint64), 3)  # fewer planes than ranks


def test_recut_moves_stay_between_neighbours():
    """SlabDecomposition::recut hands particles to ADJACENT ranks only: a cut may move at most to the far end of a
    neighbouring slab, cuts stay strictly increasing, the end cuts are fixed, and an admissible target is taken as is."""
    from sphinxsys_b200 import host
    old = np.array([0, 10, 20, 30, 40], dtype=np.int32)
    # admissible target: unchanged
    want = np.array([0, 12, 19, 33, 40], dtype=np.int32)
    assert np.array_equal(host.limit_cut_moves(old, want), want)
    # far targets are clamped into (old[r-1], old[r+1])
    got = host.limit_cut_moves(old, np.array([0, 35, 36, 37, 40], dtype=np.int32))
    assert got[0] == 0 and got[-1] == 40 and np.all(np.diff(got) >= 1)
    assert all(old[r - 1] + 1 <= got[r] <= old[r + 1] - 1 for r in range(1, 4))
    got = host.limit_cut_moves(old, np.array([0, 1, 2, 3, 40], dtype=np.int32))
    assert np.all(np.diff(got) >= 1) and all(old[r - 1] + 1 <= got[r] <= old[r + 1] - 1 for r in range(1, 4))
    # repeated re-balancing converges to the target
    cur, target = old.copy(), np.array([0, 31, 34, 37, 40], dtype=np.int32)
    for _ in range(6):
        cur = host.limit_cut_moves(cur, target)
    assert np.array_equal(cur, target)
    # random histograms: plan + limit keeps every invariant
    rng = np.random.default_rng(3)
    for _ in range(50):
        n = int(rng.integers(2, 9))
        planes = int(rng.integers(n * 3, 200))
        a = host.plan_slab_cuts(rng.integers(0, 1000, planes).astype(np.uint64) + 1, n)
        b = host.plan_slab_cuts(rng.integers(0, 1000, planes).astype(np.uint64) + 1, n)
        c = host.limit_cut_moves(a, b)
        assert c[0] == 0 and c[-1] == planes and np.all(np.diff(c) >= 1)
        assert all(a[r - 1] + 1 <= c[r] <= a[r + 1] - 1 for r in range(1, n))
