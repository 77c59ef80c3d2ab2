"""Slab-decomposed run on N GPUs against the single-GPU run of the same case (launched by torchrun, one rank per GPU).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
        tests/multi_gpu_check.py [--dp 0.05] [--outer 8] [--out gpurun_out/multi_gpu_check.json]

Checks (rank 0): every global particle id is owned by exactly one rank after K advection steps (nothing lost or
duplicated by migration), both runs took the same number of acoustic sub-steps, and positions, velocities and
densities are IDENTICAL bit for bit: ghosts are exact copies, neighbour rows have the same order (cell, then global
id) and the time-step reductions are exact maxima.
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
    ap.add_argument("--serial-exchange", action="store_true", help="plane exchange in line with the dynamics (no overlap)")
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")  # bootstrap and result gathering only; the data path is NCCL inside libsphb200
    uid = [host.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    sim = host.DamBreakCK(None, dim=3, dp=args.dp, generate=True, device_index=local, sort_interval=0, rank=rank, nranks=world,
                          unique_id=uid[0], serial_exchange=args.serial_exchange)
    sim.initialize()
    n_ac = sim.run_outer(args.outer)
    mine = {"rid": sim.download_own("ReferenceID"), "pos": sim.download_own("Position"), "vel": sim.download_own("Velocity"),
            "rho": sim.download_own("Density"), "n_ac": n_ac, "range": sim.own_range(), "cuts": sim.cuts().tolist(),
            "energy": sim.energy(), "time": sim.physical_time}
    parts = [None] * world if rank == 0 else None
    dist.gather_object(mine, parts, dst=0)
    ok, report = True, {}
    if rank == 0:
        ref = host.DamBreakCK(None, dim=3, dp=args.dp, generate=True, device_index=local, sort_interval=0)
        ref.initialize()
        n_ref = ref.run_outer(args.outer)
        n = ref.n_fluid
        rid = np.concatenate([p["rid"] for p in parts])
        report["n_global"], report["n_owned_total"] = int(n), int(rid.size)
        report["owned_per_rank"] = [int(p["rid"].size) for p in parts]
        report["stored_per_rank"] = [int(p["range"][2]) for p in parts]
        report["cuts"] = parts[0]["cuts"]
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
            ok &= report[f"bitwise_equal_{name}"]
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
    main()
