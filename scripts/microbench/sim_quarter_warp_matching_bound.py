"""Bound on what a layout matched ACROSS the eight lanes of a quarter warp could align: a row is conflict free when
the eight records read have distinct bank classes (slot mod 8); an edge colouring of lanes x classes needs max(rows, D_a) colours,
D_a = entries of class a over the eight lanes, so sum_a max(D_a - rows, 0) entries conflict whatever the layout. Lanes of a
quarter warp are storage neighbours and share their neighbourhood: their class histograms are correlated, and the bound is not
far below what the per-lane layout leaves unaligned. Same particle set as sim_local_group_alignment.py."""
import numpy as np
from scipy.spatial import cKDTree
dp = 0.0125
nx, ny, nz = 60, int(1.0 / dp), int(0.5 / dp)
rng = np.random.default_rng(0)
X = np.stack(np.meshgrid((np.arange(nx) + .5) * dp, (np.arange(ny) + .5) * dp, (np.arange(nz) + .5) * dp, indexing='ij'), -1).reshape(-1, 3)
for jitter in (0.0, 0.15):
    Y = X + jitter * dp * rng.uniform(-1, 1, X.shape)
    h = 1.3 * dp; rc = 2 * h; lower = -4 * dp - 2 * rc
    c = np.floor((Y - lower) / rc).astype(np.int64); ncell = c.max(0) + 2
    lin = (c[:, 0] * ncell[1] + c[:, 1]) * ncell[2] + c[:, 2]
    order = np.argsort(lin, kind='stable'); Ys = Y[order]
    tree = cKDTree(Ys)
    nw = len(Ys) // 32
    sel = rng.choice(np.arange(nw // 4, 3 * nw // 4), size=300, replace=False)
    forced = total = per_lane_left = 0
    for w in sel:
        slots = np.arange(w * 32, w * 32 + 32)
        nb = tree.query_ball_point(Ys[slots], rc * (1 - 1e-9))
        lists = [np.array(sorted(j for j in l if j != s)) for l, s in zip(nb, slots)]
        for q in range(4):
            ls = lists[8 * q:8 * q + 8]
            rows = max(len(l) for l in ls)
            D = np.zeros(8, dtype=np.int64)
            for l in ls:
                D += np.bincount(l % 8, minlength=8)
                hist = np.bincount(l % 8, minlength=8)
                cap = np.array([len(range(a, len(l), 8)) for a in range(8)])  # positions of each class a lane's own rows offer
                per_lane_left += np.maximum(hist - np.sort(cap)[::-1][np.argsort(np.argsort(-hist))], 0).sum()
            forced += np.maximum(D - rows, 0).sum()
            total += D.sum()
    print(f"jitter {jitter}: forced conflicts of ANY quarter-warp layout >= {forced / total:.3f} of the entries; "
          f"a lane's own histogram leaves >= {per_lane_left / total:.3f} unaligned")
