#!/usr/bin/env python
"""Hash of the particle state after a few advection steps (bit-level regression check between two builds)."""
import hashlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sphinxsys_b200.host import DamBreakCK
dp = float(sys.argv[1]) if len(sys.argv) > 1 else 0.0125
s = DamBreakCK(None, dim=3, dp=dp, fused_time_step=True, generate=True)
s.initialize()
n = s.run_outer(6)
h = hashlib.sha256()
for nm in ("Position", "Velocity", "Density"):
    h.update(s.download(nm).tobytes())
print("STATE_HASH", n, h.hexdigest()[:16], flush=True)
