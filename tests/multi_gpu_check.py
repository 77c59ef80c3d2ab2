"""Slab-decomposed run on N GPUs against the single-GPU run of the same case (launched by torchrun, one rank per GPU).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
        tests/multi_gpu_check.py [--dp 0.05] [--outer 8] [--out gpurun_out/multi_gpu_check.json]

Checks (rank 0): every global particle id is owned by exactly one rank after K advection steps (nothing lost or
duplicated by migration), both runs took the same number of acoustic sub-steps and reached the same physical time, and
positions, velocities and densities are IDENTICAL bit for bit: ghosts are exact copies, neighbour rows have the same
order (cell, then global id; bank-aligned rows are laid out relative to the slab's slot origin) and the time-step
reductions are exact maxima. `--tolerance` relaxes the pass criterion to the fp32 summation-order tolerance SURVEY.md
§8e asks for (1e-5 of the field maximum); `--trace` compares every registered variable after every advection step and
reports where the first bit differs (that is how the slot-origin dependence of the contact rows was found).
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from sphinxsys_b200 import host

    ap = argparse.ArgumentParser()
    ap.add_argument("--dp", type=float, default=0.05)
    ap.add_argument("--outer", type=int, default=8)
    ap.add_argument("--out", default="")
    ap.add_argument("--recut-interval", type=int, default=5, help="re-balance the slab cuts every so many advection steps")
    ap.add_argument("--cut-shift", type=int, default=0, help="start with the interior cuts moved by so many planes (the re-cuts undo it)")
    ap.add_argument("--strict", action="store_true", help="(default) pass only if the fields are bit-identical to the single-GPU run")
    ap.add_argument("--tolerance", action="store_true", help="pass if the fields agree within 1e-5 of the field maximum (SURVEY §8e) instead")
    ap.add_argument("--trace", action="store_true", help="compare after EVERY advection step and report where the first difference appears")
    ap.add_argument("--serial-exchange", action="store_true", help="plane exchange in line with the dynamics (no overlap)")
    ap.add_argument("--observers", action="store_true", help="the six wall-pressure probes, recorded by the rank that owns each probe's plane; not yet run on GPUs")
    ap.add_argument("--surface-indicator", action="store_true", help="FreeSurfaceIndicationCK in the loop (two sweeps around a refresh of PositionDivergence); not yet run on GPUs")
    ap.add_argument("--correction", action="store_true", help="LinearCorrectionCK variants (one more ghost refresh: the B matrix); not yet run on GPUs")
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")  # bootstrap and result gathering only; the data path is NCCL inside libsphb200
    uid = [host.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    sim = host.DamBreakCK(None, dim=3, dp=args.dp, generate=True, device_index=local, sort_interval=0, correction=args.correction, surface_indicator=args.surface_indicator, observers=args.observers, rank=rank, nranks=world,
                          unique_id=uid[0], serial_exchange=args.serial_exchange, recut_interval=args.recut_interval, initial_cut_shift=args.cut_shift)
    sim.initialize()
    cuts0 = sim.cuts().tolist()
    if args.trace:
        ref, prev_owner = None, None
        if rank == 0:
            ref = host.DamBreakCK(None, dim=3, dp=args.dp, generate=True, device_index=local, sort_interval=0, correction=args.correction, surface_indicator=args.surface_indicator, observers=args.observers)
            ref.initialize()
        for step in range(1, args.outer + 1):
            sim.run_outer(1)
            TRACED = (("rid", "ReferenceID"), ("pos", "Position"), ("vel", "Velocity"), ("rho", "Density"), ("force", "Force"),
                      ("fprior", "ForcePrior"), ("mass", "Mass"), ("C", "Compression"), ("Cdot", "CompressionRate"),
                      ("vol", "VolumetricMeasure"), ("volref", "VolumetricMeasureRef"), ("p", "Pressure"), ("dpos", "Displacement"))
            mine = {k: sim.download_own(v) for k, v in TRACED}
            mine["cuts"], mine["range"] = sim.cuts().tolist(), sim.own_range()
            mine["rel"] = [sim.exec("inner_max_count"), sim.exec("inner_stride")]
            parts = [None] * world if rank == 0 else None
            dist.gather_object(mine, parts, dst=0)
            stop = [False]
            if rank == 0:
                ref.run_outer(1)
                rid = np.concatenate([p["rid"] for p in parts])
                owner = np.concatenate([np.full(p["rid"].size, r) for r, p in enumerate(parts)])
                print("MULTI_GPU_REL", step, [p["rel"] for p in parts], [ref.exec("inner_max_count"), ref.exec("inner_stride")], flush=True)
                bad = {}
                for key, name in TRACED[1:]:
                    single = ref.download(name)
                    glob = np.zeros_like(single)
                    glob[rid] = np.concatenate([p[key] for p in parts])
                    ne = glob.view(np.uint32) != single.view(np.uint32)
                    bad[name] = np.flatnonzero(ne.reshape(single.shape[0], -1).any(axis=1))
                po = np.zeros(rid.size, dtype=np.int64)
                po[rid] = owner
                if prev_owner is None:
                    prev_owner = []
                prev_owner.append(po)
                if any(b.size for b in bad.values()):
                    x = ref.download("Position")[:, 0]
                    own_of = np.zeros(rid.size, dtype=np.int64)
                    own_of[rid] = owner
                    rep_moved = []
                    for back in (3, 2, 1):  # owner changes in the rebuilds that ended steps s-2, s-1, s
                        if len(prev_owner) > back:
                            a_, b_ = prev_owner[-back - 1], prev_owner[-back]
                            moved = np.flatnonzero(a_ != b_)
                            rep_moved.append({"rebuild_of_step": step - back + 1, "count": int(moved.size), "ids": moved[:16].tolist(),
                                              "x": [round(float(v), 4) for v in x[moved[:16]]], "from": a_[moved[:16]].tolist(), "to": b_[moved[:16]].tolist()})
                    rep = {"first_bad_step": step, "cuts": parts[0]["cuts"], "ranges": [list(map(int, p["range"])) for p in parts],
                           "changed_owner_in_previous_rebuild": rep_moved}
                    for name, b in bad.items():
                        if b.size:
                            rep[name] = {"count": int(b.size), "x_min": float(x[b].min()), "x_max": float(x[b].max()),
                                         "owners": np.bincount(own_of[b], minlength=world).tolist(), "ids": b[:12].tolist()}
                    print("MULTI_GPU_TRACE " + json.dumps(rep), flush=True)
                    if args.out:
                        json.dump(rep, open(args.out, "w"), indent=1)
                    stop[0] = True
            dist.broadcast_object_list(stop, src=0)
            if stop[0]:
                break
        else:
            if rank == 0:
                print("MULTI_GPU_TRACE " + json.dumps({"first_bad_step": None, "steps": args.outer}), flush=True)
        dist.destroy_process_group()
        sys.exit(0)
    n_ac = sim.run_outer(args.outer)
    mine = {"rid": sim.download_own("ReferenceID"), "pos": sim.download_own("Position"), "vel": sim.download_own("Velocity"),
            "rho": sim.download_own("Density"), "cuts0": cuts0, "n_ac": n_ac, "range": sim.own_range(), "cuts": sim.cuts().tolist(),
            "energy": sim.energy(), "time": sim.physical_time, "recuts": sim.exec("recuts"),
            "wall_stored": int(sim.lib.sphck_count(sim._h, 1)), "wall_global": int(sim.exec("wall_global_particles")),
            "wall_loads": int(sim.exec("wall_slab_loads")), "host_syncs": int(sim.exec("rebuild_host_syncs"))}
    if args.observers:
        mine["probes"] = sim.probe_records()[1]
    parts = [None] * world if rank == 0 else None
    dist.gather_object(mine, parts, dst=0)
    ok, report = True, {}
    if rank == 0:
        ref = host.DamBreakCK(None, dim=3, dp=args.dp, generate=True, device_index=local, sort_interval=0, correction=args.correction, surface_indicator=args.surface_indicator, observers=args.observers)
        ref.initialize()
        n_ref = ref.run_outer(args.outer)
        n = ref.n_fluid
        rid = np.concatenate([p["rid"] for p in parts])
        report["n_global"], report["n_owned_total"] = int(n), int(rid.size)
        report["owned_per_rank"] = [int(p["rid"].size) for p in parts]
        report["stored_per_rank"] = [int(p["range"][2]) for p in parts]
        report["cuts"] = parts[0]["cuts"]
        report["initial_cuts"] = parts[0]["cuts0"]
        report["recuts"] = int(parts[0]["recuts"])
        # slabs of the wall: what every rank stores of the static wall and how often re-cuts made it reload its planes
        report["wall_global"], report["wall_stored_per_rank"] = parts[0]["wall_global"], [p["wall_stored"] for p in parts]
        report["wall_slab_loads_per_rank"] = [p["wall_loads"] for p in parts]
        report["rebuild_host_round_trips_per_rank"] = [p["host_syncs"] for p in parts]
        report["acoustic_steps"] = [int(p["n_ac"]) for p in parts] + [int(n_ref)]
        ok &= rid.size == n and np.array_equal(np.sort(rid), np.arange(n, dtype=np.uint32))
        ok &= all(p["n_ac"] == n_ref for p in parts)
        for key, name, w in (("pos", "Position", 3), ("vel", "Velocity", 3), ("rho", "Density", 1)):
            glob = np.zeros((n, 3) if w == 3 else (n,), dtype=np.float32)
            glob[rid] = np.concatenate([p[key] for p in parts])
            single = ref.download(name)
            diff = float(np.max(np.abs(glob.astype(np.float64) - single.astype(np.float64)))) if n else 0.0
            report[f"max_abs_diff_{name}"] = diff
            report[f"bitwise_equal_{name}"] = bool(np.array_equal(glob.view(np.uint32), single.view(np.uint32)))
            scale = float(np.max(np.abs(single))) if n else 1.0
            report[f"rel_diff_{name}"] = diff / scale if scale else 0.0
            ok &= report[f"rel_diff_{name}"] <= 1e-5 if args.tolerance else report[f"bitwise_equal_{name}"]
        if args.observers:
            p_ref = ref.probe_records()[1]
            same = [bool(p["probes"].shape == p_ref.shape and np.array_equal(p["probes"], p_ref)) for p in parts]
            report["probe_records"], report["probe_series_equal_per_rank"] = list(p_ref.shape), same
            ok &= all(same)
        e_ref = ref.energy()
        report["energy"] = [parts[0]["energy"], e_ref]
        ok &= abs(parts[0]["energy"] - e_ref) <= 1e-9 * abs(e_ref)
        ok &= parts[0]["time"] == ref.physical_time
        report["ok"] = bool(ok)
        print("MULTI_GPU_CHECK " + json.dumps(report), flush=True)
        if args.out:
            os.makedirs(os.path.dirname(args.out), exist_ok=True)
            json.dump(report, open(args.out, "w"), indent=1)
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
