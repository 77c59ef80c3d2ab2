#!/usr/bin/env python
"""Long-horizon probe of the dam-break loop: acoustic sub-steps per advection step, max |v|, energy every `every` steps.
    python scripts/long_run_probe.py [dp] [steps] [every]            (one GPU)
    torchrun ... scripts/long_run_probe.py [dp] [steps] [every]      (decomposed: same numbers, all-reduced)"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from sphinxsys_b200 import host
dp = float(sys.argv[1]) if len(sys.argv) > 1 else 0.00625
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1300
every = int(sys.argv[3]) if len(sys.argv) > 3 else 100
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
uid = None
if world > 1:
    import torch, torch.distributed as dist
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")
    box = [host.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    uid = box[0]
s = host.DamBreakCK(None, dim=3, dp=dp, generate=True, device_index=local, sort_interval=100, recut_interval=100, rank=rank, nranks=world, unique_id=uid)
s.initialize()
done = 0
while done < steps:
    n_ac = s.run_outer(every)
    done += every
    v = s.download_own("Velocity")
    vmax = float(np.sqrt((v.astype(np.float64) ** 2).sum(axis=1)).max()) if v.size else 0.0
    x = s.download_own("Position")
    rec = {"steps": done, "t": s.physical_time, "acoustic_per_outer": n_ac / every, "vmax": vmax, "energy": s.energy(),
           "x_front": float(x[:, 0].max()) if x.size else 0.0, "y_top": float(x[:, 1].max()) if x.size else 0.0, "own": int(x.shape[0])}
    if world > 1:
        allr = [None] * world
        dist.all_gather_object(allr, rec)
        rec["vmax"] = max(r["vmax"] for r in allr); rec["x_front"] = max(r["x_front"] for r in allr); rec["y_top"] = max(r["y_top"] for r in allr)
        rec["own"] = [r["own"] for r in allr]
    if rank == 0:
        print("PROBE", json.dumps(rec), flush=True)
if rank == int(os.environ.get("TRACE_RANK", "0")):
    s.step_trace_report(done)
