#!/usr/bin/env python
"""Kernel timing harness for tuning: config 2 (or --dp), times each hot-path kernel with CUDA events on the live state.
usage: python scripts/kbench.py [--dp 0.00625] [--outer 3]   -> one JSON line"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from bench import time_kernel  # noqa: E402
from sphinxsys_b200.host import DamBreakCK  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--dp", type=float, default=0.00625)
ap.add_argument("--outer", type=int, default=3)
ap.add_argument("--tag", default="")
a = ap.parse_args()
s = DamBreakCK(None, dim=3, dp=a.dp, fused_time_step=True, generate=True)
s.initialize()
s.run_outer(a.outer)
torch.cuda.synchronize()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ac0 = s.acoustic_steps
ev0.record()
s.run_outer(5)
ev1.record()
torch.cuda.synchronize()
ms = ev0.elapsed_time(ev1) / 5
n_ac = (s.acoustic_steps - ac0) / 5
dt = s.last_acoustic_dt * 1e-3
out = {"tag": a.tag, "ms_per_outer": ms, "G_particle_steps_s": s.n_fluid * n_ac / ms / 1e6,
       "a2": time_kernel(lambda: s.exec("acoustic2", dt), 20, torch),
       "a1": time_kernel(lambda: s.exec("acoustic1", dt), 20, torch),
       "density": time_kernel(lambda: s.exec("density_summation"), 10, torch),
       "cell_list": time_kernel(lambda: s.exec("rebuild"), 10, torch),
       "relations": time_kernel(lambda: s.exec("relations"), 5, torch)}
print(json.dumps(out), flush=True)
