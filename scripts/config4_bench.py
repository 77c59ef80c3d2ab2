#!/usr/bin/env python
"""BASELINE config 4: periodic Taylor-Green vortex, n_side^3 particles PER GPU (default 256^3 = 16,777,216), weak scaling.

One GPU (plain python): two spellings of the periodic condition — images on all three axes (TaylorGreenCK) and the
ring of ONE slab (x periodic through the slab exchange + seam shift, images on y / z: what every rank of an N-GPU run
does, minus the NCCL transport).
N GPUs (torchrun, one rank per GPU): the box replicated along x, [0, N] x [0, 1]^2, as a ring of N slabs; every rank
generates its own copy of the unit box only.

    python scripts/config4_bench.py [n_side] [outer steps] [modes: ring,images]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 \
        scripts/config4_bench.py 256 6

Wall clock around run_outer (host round trips of the configuration update included), max over ranks."""
import dataclasses
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from sphinxsys_b200 import cases, host  # noqa: E402

n_side = int(sys.argv[1]) if len(sys.argv) > 1 else 256
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
modes = (sys.argv[3] if len(sys.argv) > 3 else "ring,images").split(",")
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
dist = None
uid = None
if world > 1:
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")  # bootstrap and timing only; the data path is NCCL inside libsphb200
    box = [host.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    uid, modes = box[0], ["ring"]

t0 = time.perf_counter()
case = cases.taylor_green(dim=3, n_side=n_side, jitter=0.05)
n_local = case.n_fluid
local_ids = None
if world > 1:
    # this rank's copy of the unit box at x in [rank, rank + 1): same lattice, same jitter, velocity field of period 1
    pos = case.fluid_pos.copy()
    pos[:, 0] += np.float32(rank)
    up = list(case.periodic_upper)
    up[0] = float(world)
    case = dataclasses.replace(case, fluid_pos=pos, DL=float(world), LL=float(world), periodic_upper=tuple(up))
    local_ids = np.arange(n_local, dtype=np.uint32) + np.uint32(rank * n_local)
if rank == 0:
    print(f"case {n_local} particles per rank x {world} in {time.perf_counter() - t0:.1f} s", flush=True)
out = {"n_side": n_side, "particles_per_gpu": n_local, "gpus": world, "outer_steps": steps}
for mode in modes:
    t0 = time.perf_counter()
    if mode == "ring":
        gpu = host.TaylorGreenCK(case, device_index=local, ring=True, rank=rank, nranks=world, unique_id=uid, local_ids=local_ids,
                                 sort_interval=0)
    else:
        gpu = host.TaylorGreenCK(case, device_index=local, sort_interval=0)
    gpu.initialize()
    gpu.run_outer(2)  # warm-up
    gpu.synchronize()
    l0 = gpu.launches
    if dist:
        dist.barrier()
    t1 = time.perf_counter()
    n_ac = gpu.run_outer(steps)
    gpu.synchronize()
    dt = time.perf_counter() - t1
    if rank == 0:
        gpu.step_trace_report(steps + 2)  # SPHB200_STEP_TRACE=1 only
    if dist:
        every = [None] * world
        dist.all_gather_object(every, dt)
        dt = max(every)
    rec = {"setup_s": t1 - t0, "ms_per_outer_step": 1e3 * dt / steps, "acoustic_steps": n_ac,
           "particle_steps_per_s": world * n_local * n_ac / dt, "launches": gpu.launches - l0,
           "images": gpu.ghost_particles, "plane_ghosts": int(gpu.exec("plane_ghost_particles")) if mode == "ring" else 0,
           "energy": gpu.energy()}
    out[mode] = rec
    if rank == 0:
        print(mode, json.dumps(rec), flush=True)
    gpu.close()
    del gpu
if rank == 0:
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"config4_{n_side}_{world}gpu.json"), "w"), indent=1)
if dist:
    dist.destroy_process_group()
