#!/usr/bin/env python
"""BASELINE config 4 on one GPU: periodic Taylor-Green vortex, n_side^3 particles (default 256^3 = 16,777,216).
Two spellings of the periodic condition: images on all three axes (TaylorGreenCK) and the ring of ONE slab (x periodic
through the slab exchange + seam shift, images on y / z: what every rank of an N-GPU run does, minus the NCCL transport).
usage: scripts/config4_bench.py [n_side] [outer steps] [modes: ring,images]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from sphinxsys_b200 import cases  # noqa: E402
from sphinxsys_b200.host import TaylorGreenCK  # noqa: E402

n_side = int(sys.argv[1]) if len(sys.argv) > 1 else 256
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
modes = (sys.argv[3] if len(sys.argv) > 3 else "ring,images").split(",")
t0 = time.perf_counter()
case = cases.taylor_green(dim=3, n_side=n_side, jitter=0.05)
print(f"case {case.n_fluid} particles in {time.perf_counter() - t0:.1f} s", flush=True)
out = {"n_side": n_side, "particles": case.n_fluid, "outer_steps": steps}
for mode in modes:
    t0 = time.perf_counter()
    gpu = TaylorGreenCK(case, ring=(mode == "ring"), sort_interval=0)
    gpu.initialize()
    gpu.run_outer(2)  # warm-up
    gpu.synchronize()
    l0 = gpu.launches
    t1 = time.perf_counter()
    n_ac = gpu.run_outer(steps)
    gpu.synchronize()
    dt = time.perf_counter() - t1
    rec = {"setup_s": t1 - t0, "ms_per_outer_step": 1e3 * dt / steps, "acoustic_steps": n_ac,
           "particle_steps_per_s": case.n_fluid * n_ac / dt, "launches": gpu.launches - l0,
           "images": gpu.ghost_particles, "plane_ghosts": int(gpu.exec("plane_ghost_particles")) if mode == "ring" else 0,
           "energy": gpu.energy()}
    out[mode] = rec
    print(mode, json.dumps(rec), flush=True)
    gpu.close()
    del gpu
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"config4_{n_side}.json"), "w"), indent=1)
