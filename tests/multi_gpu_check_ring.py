"""Periodic ring of slabs on N GPUs (BASELINE config 4) against the ring of ONE slab and the single-domain oracle
(launched by torchrun, one rank per GPU; NOT collected by pytest).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
        tests/multi_gpu_check_ring.py [--n-side 32] [--outer 12] [--drift 2.0] [--out gpurun_out/ring_check.json]

Rank r owns the cell planes [first + P r / N, first + P (r+1) / N) of the aligned mesh (the cuts TaylorGreenCK makes
itself) and hands its particles with their global numbers to the case. Checks (rank 0): every particle owned exactly
once after K advection steps with a drift through the seam; every rank took the acoustic steps of the reference runs;
fields against (a) the ring of one slab on this GPU — same kernels, same cell order, expected bit-identical — and (b)
the fp32 oracle's single-domain periodic run on the same mesh, within the tolerances of
tests/test_gpu_zz_periodic_ring.py. State of round 1: the ring of one slab is verified on B200, this N >= 2 run (the
NCCL transport, same-peer ordering at N = 2) has not been executed yet.
"""
import argparse
import dataclasses
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist
    from helpers import rel_err
    from oracle import decomposed as dec
    from oracle import oracle as orc
    from sphinxsys_b200 import cases, host, hostmath as hm

    ap = argparse.ArgumentParser()
    ap.add_argument("--n-side", type=int, default=32)
    ap.add_argument("--outer", type=int, default=12)
    ap.add_argument("--drift", type=float, default=2.0)
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")  # bootstrap and result gathering only; the data path is NCCL inside libsphb200
    uid = [host.comm_unique_id() if rank == 0 and world > 1 else None]
    dist.broadcast_object_list(uid, src=0)

    case = cases.taylor_green(dim=3, n_side=args.n_side, jitter=0.05)
    vel = case.fluid_vel.copy()
    vel[:, 0] += np.float32(args.drift)
    case = dataclasses.replace(case, fluid_vel=vel)
    m, sm = host.aligned_periodic_mesh(case.periodic_lower, case.periodic_upper, case.kernel.cutoff, case.dim)
    mesh = hm.MeshSpec(tuple(float(v) for v in m.lower), float(m.spacing), tuple(int(c) for c in m.cells))
    cuts = [sm.first_plane + sm.box_planes * r // world for r in range(world + 1)]
    plane = dec.x_plane(case.fluid_pos, mesh)
    own = np.flatnonzero((plane >= cuts[rank]) & (plane < cuts[rank + 1]))

    sim = host.TaylorGreenCK(case, device_index=local, ring=True, rank=rank, nranks=world, unique_id=uid[0], own=own)
    sim.initialize()
    n_ac = sim.run_outer(args.outer)
    mine = {"rid": sim.download_own("ReferenceID"), "acoustic": n_ac, "range": sim.own_range(),
            "plane_ghosts": int(sim.exec("plane_ghost_particles")), "images": sim.ghost_particles}
    for nm in ("Position", "Velocity", "Density", "Compression"):
        mine[nm] = sim.download_own(nm)
    energy = sim.energy()  # collective (all-reduce): every rank calls it
    parts = [None] * world if rank == 0 else None
    dist.gather_object(mine, parts, dst=0)
    sim.close()
    ok = True
    if rank == 0:
        n = case.n_fluid
        rid = np.concatenate([p["rid"] for p in parts]).astype(np.int64)
        rep = {"ranks": world, "n_side": args.n_side, "outer": args.outer, "cuts": cuts, "own": [int(p["rid"].size) for p in parts],
               "plane_ghosts": [p["plane_ghosts"] for p in parts], "images": [p["images"] for p in parts],
               "acoustic": [p["acoustic"] for p in parts], "energy": energy}
        rep["partition"] = bool(rid.size == n and np.array_equal(np.sort(rid), np.arange(n)))
        ok &= rep["partition"]
        glob = {}
        if rep["partition"]:
            for nm in ("Position", "Velocity", "Density", "Compression"):
                a = np.concatenate([p[nm] for p in parts])
                g = np.empty_like(a)
                g[rid] = a
                glob[nm] = g
            # (a) ring of one slab on this GPU
            one = host.TaylorGreenCK(case, device_index=local, ring=True)
            one.initialize()
            ac_one = one.run_outer(args.outer)
            r1 = one.download_own("ReferenceID").astype(np.int64)
            rep["vs_ring_of_one"] = {"acoustic": ac_one}
            for nm in glob:
                a = one.download_own(nm)
                g = np.empty_like(a)
                g[r1] = a
                rep["vs_ring_of_one"][nm] = {"bit_identical": bool(np.array_equal(g.view(np.uint32), glob[nm].view(np.uint32))),
                                             "rel_err": rel_err(glob[nm], g)}
            rep["vs_ring_of_one"]["energy"] = one.energy()
            one.close()
            ok &= all(p["acoustic"] == ac_one for p in parts)
            # (b) single-domain oracle on the same mesh
            o32 = orc.OracleSim(dataclasses.replace(case, mesh=mesh), free_surface=0)
            o32.exec("prepare_ck")
            o32.exec("run_ck", 1e9, args.outer, 1e9, 0)
            rep["vs_oracle"] = {"acoustic": int(o32.exec("acoustic_steps")), "energy": o32.exec("energy")}
            d = glob["Position"].astype(np.float64) - o32.real("Position", 3).reshape(-1, 3)
            d -= np.round(d)
            rep["vs_oracle"]["Position"] = float(np.abs(d).max())
            ok &= rep["vs_oracle"]["Position"] < 5e-6
            for nm, w, tol in (("Velocity", 3, 2e-4), ("Density", 1, 2e-6), ("Compression", 1, 2e-6)):
                b = o32.real(nm, w).reshape(-1, w) if w > 1 else o32.real(nm, w)
                rep["vs_oracle"][nm] = rel_err(glob[nm], b)
                ok &= rep["vs_oracle"][nm] < tol
            ok &= abs(energy - rep["vs_oracle"]["energy"]) <= 1e-5 * abs(rep["vs_oracle"]["energy"])
            rep["crossed_the_seam"] = int((np.abs(glob["Position"][:, 0].astype(np.float64) - case.fluid_pos[:, 0]) > 0.5).sum())
        rep["ok"] = bool(ok)
        print("RING_CHECK", json.dumps(rep, default=float), flush=True)
        if args.out:
            os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
            json.dump(rep, open(args.out, "w"), indent=1, default=float)
    flag = [ok]
    dist.broadcast_object_list(flag, src=0)
    dist.destroy_process_group()
    sys.exit(0 if flag[0] else 1)


if __name__ == "__main__":
    try:
        main()
    except SystemExit:
        raise
    except BaseException:  # a rank-local failure must not leave the peers waiting in a collective: die at once, loudly
        import traceback
        traceback.print_exc()
        sys.stderr.flush()
        os._exit(1)
