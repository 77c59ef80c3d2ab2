#!/usr/bin/env python
"""The complete reference case file (correction aliases + free-surface indication + probes) for a few advection steps:
run under `ncu --metrics gpu__time_duration.sum` for the launch list, or alone for ms per step.
usage: python scripts/complete_case_probe.py [--dp 0.00625] [--outer 3]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from sphinxsys_b200.host import DamBreakCK  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--dp", type=float, default=0.00625)
ap.add_argument("--outer", type=int, default=3)
ap.add_argument("--warmup", type=int, default=2)
a = ap.parse_args()
s = DamBreakCK(None, dim=3, dp=a.dp, fused_time_step=True, sort_interval=100, generate=True, correction=True, surface_indicator=True,
               observers=True)
s.initialize()
s.run_outer(a.warmup)
torch.cuda.synchronize()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ac0 = s.acoustic_steps
ev0.record()
s.run_outer(a.outer)
ev1.record()
torch.cuda.synchronize()
ms = ev0.elapsed_time(ev1) / a.outer
print("COMPLETE_CASE", json.dumps({"dp": a.dp, "n_fluid": int(s.n_fluid), "ms_per_step": ms, "acoustic_steps_per_outer": (s.acoustic_steps - ac0) / a.outer,
                                   "energy": s.energy(), "tile_order": os.environ.get("SPHB200_TILE_ORDER", "default")}), flush=True)
