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


@pytest.mark.parametrize("n_side", [16, 20, 32, 100, 256])
def test_ring_seam_thresholds_follow_the_cell_arithmetic(n_side):
    """Periodic ring of slabs (config 4 on N GPUs): the cell planes tile the box and the seam thresholds handed to
    sphb200_seam_shift are exactly the x ranges of the planes in the mesh's own cell arithmetic (oracle cell keys), so a
    shifted particle can be held in the plane it belongs to: leavers inside the box planes, boundary planes in the one
    ghost plane beyond the face."""
    from oracle import oracle as orc
    from sphinxsys_b200 import host, hostmath as hm
    f = np.float32
    cutoff = f(2.0) * f(1.3) * f(1.0 / n_side)
    m, sm = host.aligned_periodic_mesh((0, 0, 0), (1, 1, 1), cutoff)
    mesh = hm.MeshSpec(tuple(m.lower), m.spacing, tuple(m.cells))
    k0, P = sm.first_plane, sm.box_planes
    assert m.spacing >= cutoff and P == int(1.0 / cutoff) and m.cells[0] == P + 2 * k0

    def plane(x):
        pos = np.full((len(x), 3), 0.5, dtype=f)
        pos[:, 0] = x
        cell, _ = orc.cell_keys(pos, mesh)
        return cell.astype(np.int64) // (m.cells[1] * m.cells[2])

    t = np.array([sm.ghost_low_min, sm.ghost_low_max, sm.own_min, sm.own_max, sm.ghost_high_min, sm.ghost_high_max], dtype=f)
    assert plane(t).tolist() == [k0 - 1, k0 - 1, k0, k0 + P - 1, k0 + P, k0 + P]
    beyond = np.array([np.nextafter(t[0], f(-9)), np.nextafter(t[2], f(-9)), np.nextafter(t[3], f(9)), np.nextafter(t[5], f(9))], dtype=f)
    assert plane(beyond).tolist() == [k0 - 2, k0 - 1, k0 + P, k0 + P + 1]
    assert abs(float(sm.own_min)) < 1e-7 and abs(float(sm.ghost_high_min) - 1.0) < 1e-6

    # k_seam_shift restated: what crosses the seam lands in the plane its sender saw it in, moved by +/- P planes
    rng = np.random.default_rng(n_side)
    s, L = f(m.spacing), f(1.0)
    for delta in (L, -L):
        if delta > 0:   # rank 0 -> last rank: its first plane and its leavers (just below the box)
            x = np.concatenate([rng.uniform(-0.3 * s, 0.0, 4000), rng.uniform(0.0, s, 4000), [sm.own_min, np.nextafter(f(sm.own_min), f(-9)),
                                np.nextafter(f(sm.own_min) + s, f(-9)), -1e-30, 0.0]]).astype(f)
            x = x[plane(x) <= k0]
            y = x + delta
            leaver = plane(x) < k0
            y = np.where(leaver, np.minimum(y, f(sm.own_max)), np.minimum(np.maximum(y, f(sm.ghost_high_min)), f(sm.ghost_high_max)))
        else:           # last rank -> rank 0: its last plane and its leavers (just above the box)
            x = np.concatenate([rng.uniform(1.0, 1.0 + 0.3 * s, 4000), rng.uniform(1.0 - s, 1.0, 4000),
                                [sm.own_max, sm.ghost_high_min, 1.0, np.nextafter(f(1.0), f(9))]]).astype(f)
            x = x[plane(x) >= k0 + P - 1]
            y = x + delta
            leaver = plane(x) >= k0 + P
            y = np.where(leaver, np.maximum(y, f(sm.own_min)), np.maximum(np.minimum(y, f(sm.ghost_low_max)), f(sm.ghost_low_min)))
        moved = plane(y.astype(f)) - plane(x)
        assert np.all(moved == (P if delta > 0 else -P)), "a shifted particle changed its plane relative to the box"
        free = (y == x + delta)
        assert free.mean() > 0.99, "the clamps act in rounding cases only"
