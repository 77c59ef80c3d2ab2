#!/usr/bin/env python
"""Where the e2e step loses against the device-resident step (one GPU, config 2): the pipelined step with parts switched off."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from sphinxsys_b200 import host
from sphinxsys_b200.host import DamBreakCK
dp = float(sys.argv[1]) if len(sys.argv) > 1 else 0.00625
K = 12
s = DamBreakCK(None, dim=3, dp=dp, fused_time_step=True, sort_interval=0, generate=True)
s.initialize()
s.run_outer(5)
n = s.n_fluid
in_names = ["Position", "VolumetricMeasure", "Velocity", "Mass", "ForcePrior", "Compression", "CompressionRate", "VolumetricMeasureRef", "PreviousGravityForceCK"]
out_names = ["Position", "Velocity", "Density"]
def pinned(name):
    w = 3 if name in host.VEC_NAMES else 1
    t = torch.empty((n, w) if w > 1 else (n,), dtype=torch.float32).pin_memory()
    return t, t.numpy()
hin = {nm: pinned(nm) for nm in in_names}
hout = {nm: pinned(nm) for nm in out_names}
for nm in in_names:
    s.download(nm, out=hin[nm][1])
ins = [hin[nm][1] for nm in in_names]
outs = [hout[nm][1] for nm in out_names]
def timed(fn, label):
    fn(2)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    fn(K)
    torch.cuda.synchronize()
    ms = 1e3 * (time.perf_counter() - t0) / K
    print("E2E_PROBE", json.dumps({"variant": label, "ms_per_step": round(ms, 3)}), flush=True)
timed(lambda k: s.run_outer(k), "device-resident loop (configuration update after the dynamics)")
s.exec("configuration_before_dynamics", 1.0)
timed(lambda k: [s.step_outer() for _ in range(k)], "configuration update first, first acoustic dt unfused, no copies")
s.pipeline_create(in_names, out_names)
def both(k, up=True, down=True):
    if up:
        s.pipeline_stage_uploads(ins)
    for i in range(k):
        if up:
            s.pipeline_commit_uploads()
            if i + 1 < k:
                s.pipeline_stage_uploads(ins)
        s.step_outer()
        if down:
            s.pipeline_stage_downloads(outs)
    s.pipeline_synchronize()
timed(lambda k: both(k, True, False), "+ uploads (H2D on the side stream, commit on the main stream)")
timed(lambda k: both(k, False, True), "+ downloads only")
timed(lambda k: both(k, True, True), "+ uploads and downloads (the e2e step)")
