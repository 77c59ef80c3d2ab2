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


@pytest.mark.parametrize("n_side,nranks", [(16, 1), (16, 2), (32, 3), (256, 8)])
def test_ring_ownership_never_loses_or_duplicates_a_particle(n_side, nranks):
    """The ownership rules of the GPU ring restated in numpy (SlabDecomposition::rebuild on a ring: leavers and boundary
    planes selected by cell PLANE, sphb200_seam_shift on what crosses the seam) and driven hard: particles start within a
    few ulps of the box faces and of the cuts and random-walk by up to a fifth of a plane per step. After every
    configuration update each particle has exactly one owner, sits in one of its owner's planes, and every rank's ghost
    planes hold exactly the particles of its neighbours' boundary planes (sizes that do not match would leave an NCCL
    receive waiting: the invariant refreshGhosts() relies on)."""
    from oracle import oracle as orc
    from sphinxsys_b200 import host, hostmath as hm
    f = np.float32
    cutoff = f(2.0) * f(1.3) * f(1.0 / n_side)
    m, sm = host.aligned_periodic_mesh((0, 0, 0), (1, 1, 1), cutoff)
    mesh = hm.MeshSpec(tuple(m.lower), m.spacing, tuple(m.cells))
    k0, P, s, L = sm.first_plane, sm.box_planes, f(m.spacing), f(1.0)
    cuts = [k0 + P * r // nranks for r in range(nranks + 1)]

    def plane(x):
        if len(x) == 0:
            return np.zeros(0, dtype=np.int64)
        pos = np.full((len(x), 3), 0.5, dtype=f)
        pos[:, 0] = x
        cell, _ = orc.cell_keys(pos, mesh)
        return cell.astype(np.int64) // (m.cells[1] * m.cells[2])

    def shift(x, delta):  # k_seam_shift
        y = (x + delta).astype(f)
        pl = plane(x)
        if delta > 0:
            return np.where(pl < k0, np.minimum(y, f(sm.own_max)), np.minimum(np.maximum(y, f(sm.ghost_high_min)), f(sm.ghost_high_max))).astype(f)
        return np.where(pl >= k0 + P, np.maximum(y, f(sm.own_min)), np.maximum(np.minimum(y, f(sm.ghost_low_max)), f(sm.ghost_low_min))).astype(f)

    rng = np.random.default_rng(n_side + nranks)
    faces = np.array([0.0, 1.0] + [float(m.lower[0]) + c * float(s) for c in cuts[1:-1]])
    n = 4000
    x0 = (rng.choice(faces, n) + rng.choice([0.0, 1e-9, -1e-9, 6e-8, -6e-8, 1e-4, -1e-4], n) + np.where(rng.random(n) < 0.3, rng.uniform(0, 1, n), 0.0))
    x0 = np.mod(x0, 1.0).astype(f)
    x0 = np.clip(x0, f(sm.own_min), f(sm.own_max))
    pl0 = plane(x0)
    own = [{"gid": np.flatnonzero((pl0 >= cuts[r]) & (pl0 < cuts[r + 1])), "x": None} for r in range(nranks)]
    for r in range(nranks):
        own[r]["x"] = x0[own[r]["gid"]]
    assert sum(o["gid"].size for o in own) == n
    for step in range(30):
        send = []
        for r in range(nranks):  # move, then sphb200_slab_select: <= first own plane to the left, >= last own plane to the right
            o = own[r]
            o["x"] = (o["x"] + rng.uniform(-0.2, 0.2, o["x"].size).astype(f) * s).astype(f)
            # the rounding cases on purpose: some particles of the seam ranks are put within a few ulps of the box faces,
            # on either side (just outside: leavers whose shifted position rounds onto the far face; just inside: boundary-
            # plane particles whose image rounds into the box)
            pl = plane(o["x"])
            if r == 0:
                near = np.flatnonzero(pl == cuts[0])[:24]
                edge = f(sm.own_min)
                vals = [edge, np.nextafter(edge, f(9)), np.nextafter(edge, f(-9)), np.nextafter(np.nextafter(edge, f(-9)), f(-9)), f(-2e-8), f(0.0)]
                o["x"][near] = np.array([vals[k % len(vals)] for k in range(near.size)], dtype=f)
            if r == nranks - 1:
                near = np.flatnonzero(pl == cuts[-1] - 1)[-24:]
                edge = f(sm.own_max)
                vals = [edge, np.nextafter(edge, f(-9)), f(sm.ghost_high_min), np.nextafter(f(sm.ghost_high_min), f(9)), f(1.0), np.nextafter(f(1.0), f(9))]
                o["x"][near] = np.array([vals[k % len(vals)] for k in range(near.size)], dtype=f)
            pl = plane(o["x"])
            send.append({"left": pl <= cuts[r], "right": pl >= cuts[r + 1] - 1})
        new = []
        for r in range(nranks):
            o, lft, rgt = own[r], (r - 1) % nranks, (r + 1) % nranks
            from_left = (own[lft]["gid"][send[lft]["right"]], own[lft]["x"][send[lft]["right"]])
            from_right = (own[rgt]["gid"][send[rgt]["left"]], own[rgt]["x"][send[rgt]["left"]])
            xl = shift(from_left[1], -L) if r == 0 else from_left[1]                  # over the seam: my left neighbour is the last rank
            xr = shift(from_right[1], L) if r == nranks - 1 else from_right[1]
            gid = np.concatenate([o["gid"], from_left[0], from_right[0]])
            x = np.concatenate([o["x"], xl, xr])
            pl = plane(x)
            mine = (pl >= cuts[r]) & (pl < cuts[r + 1])
            ghosts_l, ghosts_r = pl < cuts[r], pl >= cuts[r + 1]
            new.append({"gid": gid[mine], "x": x[mine], "gl": np.sort(gid[ghosts_l]), "gr": np.sort(gid[ghosts_r]),
                        "pl_ok": bool(np.all(pl[ghosts_l] == cuts[r] - 1) and np.all(pl[ghosts_r] == cuts[r + 1]))})
        own = new
        owners = np.concatenate([o["gid"] for o in own])
        assert owners.size == n and np.array_equal(np.sort(owners), np.arange(n)), f"step {step}: a particle was lost or is owned twice"
        assert all(o["pl_ok"] for o in own), f"step {step}: a ghost outside the ghost planes"
        for r in range(nranks):  # ghost planes = the neighbours' boundary planes after the update (a ring of one: its own)
            lft, rgt = (r - 1) % nranks, (r + 1) % nranks
            pl_l, pl_r = plane(own[lft]["x"]), plane(own[rgt]["x"])
            assert np.array_equal(own[r]["gl"], np.sort(own[lft]["gid"][pl_l == cuts[lft + 1] - 1])), f"step {step}: left ghost plane of rank {r}"
            assert np.array_equal(own[r]["gr"], np.sort(own[rgt]["gid"][pl_r == cuts[rgt]])), f"step {step}: right ghost plane of rank {r}"


def test_wall_slab_planning_invariants():
    """WallSlab::plan (include/sphinxsys_ck/dambreak_case.h): the wall planes a rank stores around its fluid planes. Whatever the
    histogram, the cuts and the storage bound: (i) the planes the contact search can reach are always stored; (ii) the margin is
    kept in full when the storage allows it and given up only as far as needed; (iii) nothing is reloaded while the stored planes
    still cover the need (re-cuts inside the margin are free); (iv) if even the needed planes exceed the bound, exactly those are
    planned (the load then fails loudly in BaseParticles, not silently here)."""
    from sphinxsys_b200 import host
    rng = np.random.default_rng(17)
    for _ in range(300):
        planes = int(rng.integers(8, 120))
        h = rng.integers(0, 500, planes).astype(np.uint64)
        h[rng.random(planes) < 0.2] = 0  # empty planes (no wall there)
        below = np.concatenate([[0], np.cumsum(h)])
        depth, margin = int(rng.integers(1, 3)), int(rng.integers(0, 6))
        X0 = int(rng.integers(0, planes - 1))
        X1 = int(rng.integers(X0 + 1, planes + 1))
        need_lo, need_hi = max(0, X0 - depth), min(planes - 1, X1 - 1 + depth)
        need = int(below[need_hi + 1] - below[need_lo])
        full_lo, full_hi = max(0, need_lo - margin), min(planes - 1, need_hi + margin)
        full = int(below[full_hi + 1] - below[full_lo])
        bound = int(rng.integers(max(need // 2, 1), full + 50))
        lo, hi, reload = host.wall_slab_plan(h, depth, margin, X0, X1, bound)
        assert reload and lo <= need_lo and hi >= need_hi                                       # (i)
        stored = int(below[hi + 1] - below[lo])
        if bound >= full:
            assert (lo, hi) == (full_lo, full_hi)                                              # (ii) whole margin
        elif bound >= need:
            assert stored <= bound and full_lo <= lo and hi <= full_hi                         # (ii) margin given up as far as needed
        else:
            assert (lo, hi) == (need_lo, need_hi)                                              # (iv)
        # (iii) a re-cut that stays inside the stored planes does not reload and leaves them as they are
        if lo + depth < hi - depth:
            Y0 = int(rng.integers(lo + depth, hi - depth + 1)) if lo > 0 else int(rng.integers(0, hi - depth + 1))
            Y1 = int(rng.integers(Y0 + 1, hi - depth + 2)) if hi < planes - 1 else int(rng.integers(Y0 + 1, planes + 1))
            if max(0, Y0 - depth) >= lo and min(planes - 1, Y1 - 1 + depth) <= hi:
                lo2, hi2, reload2 = host.wall_slab_plan(h, depth, margin, Y0, Y1, bound, stored=(lo, hi))
                assert not reload2 and (lo2, hi2) == (lo, hi)
