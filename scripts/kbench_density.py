#!/usr/bin/env python
"""Time one kernel of the live config-2 state quickly (tuning helper). usage: kbench_density.py <tag> [op]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from bench import time_kernel
from sphinxsys_b200.host import DamBreakCK
tag = sys.argv[1]
ops = sys.argv[2].split(",") if len(sys.argv) > 2 else ["density_summation"]
s = DamBreakCK(None, dim=3, dp=0.00625, fused_time_step=True, generate=True)
s.initialize()
s.run_outer(2)
torch.cuda.synchronize()
dt = s.last_acoustic_dt * 1e-3
out = {"tag": tag}
for op in ops:
    arg = dt if op.startswith("acoustic") else 0.0
    out[op] = time_kernel(lambda: s.exec(op, arg), 20, torch)
print(json.dumps(out), flush=True)
