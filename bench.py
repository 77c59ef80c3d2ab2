#!/usr/bin/env python
"""bench.py — particle-steps/s of the 3-D WCSPH dam break (BASELINE.json metric) on N B200s of one node.

One "step" = one advection (outer) step of the case-file loop (dambreak.cpp:188-222): density summation +
regularisation, advection set-up, advection-dt reduction, the ~5 acoustic sub-steps it contains (acoustic-dt
reduction + 1st half + 2nd half each), position update, particle sort at its natural cadence (every 100 outer
steps), cell-linked-list and relation rebuild.  value = N_fluid x (acoustic sub-steps executed) / seconds, i.e.
every per-advection cost is amortised into the particle-step rate.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--dp 0.00625] [--impl ours|reference]

Prints ONE JSON line (rank 0).  `--impl reference` times the CPU oracle (the reference cannot be built in this
image) on the box's host cores on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

# stdout carries exactly ONE line, the JSON: NCCL's version banner / debug log (printed when the environment sets NCCL_DEBUG)
# goes to stderr unless the caller chose a file for it
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

A_STEP_1ST = 128  # algorithmic bytes / particle, 1st half  
A_STEP_2ND = 84   # algorithmic bytes / particle, 2nd half
A_OUTER = 122     # per outer step extras
A_SORT = 224      # per sort


def measured_peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_baseline_run(dp, max_outer, budget_s, threads=0, warmup=0):
    """The oracle (CPU restatement of the reference's CK par_host data flow) on the host cores: bounded sample.
    `warmup` outer steps run untimed first; then up to `max_outer` timed outer steps (fewer only if budget_s runs out)."""
    from oracle import oracle as orc
    from sphinxsys_b200 import cases
    case = cases.dam_break(dim=3, dp=dp)
    sim = orc.OracleSim(case, f64=False, threads=threads)
    sim.exec("prepare_ck")
    for _ in range(warmup):
        sim.exec("run_ck", 1e9, 1, 1e9, 100)
    n_ac0 = int(sim.exec("acoustic_steps"))
    t0 = time.perf_counter()
    done = 0
    while done < max_outer and (time.perf_counter() - t0) < budget_s:
        sim.exec("run_ck", 1e9, 1, 1e9, 100)
        done += 1
    elapsed = time.perf_counter() - t0
    n_ac = int(sim.exec("acoustic_steps")) - n_ac0
    # also ns per pair interaction = wall time over (neighbour-list entries x 2 half steps x acoustic steps);
    # the per-advection work (summation, cell list, relation search) is inside the wall time, as in `value`
    pairs = int(sim.uint("inner_offset")[-1]) + int(sim.uint("contact_offset")[-1])
    return {"value": case.n_fluid * n_ac / elapsed, "unit": "particle-steps/s", "cores": orc.lib().orc_max_threads(),
            "kind": "port", "ns_per_pair_interaction": 1e9 * elapsed / max(2.0 * pairs * n_ac, 1.0),
            "sample": f"3-D dam break dp={dp} ({case.n_fluid} fluid + {case.n_wall} wall), {done} outer / {n_ac} acoustic steps "
                      f"in {elapsed:.1f} s after {warmup} untimed, oracle fp32 + OpenMP (restatement of the reference CK par_host "
                      f"path, not the TBB build)",
            "_elapsed": elapsed, "_steps": done, "_ms_per_step": 1e3 * elapsed / max(done, 1), "_n_fluid": case.n_fluid,
            "_n_wall": case.n_wall, "_acoustic_per_outer": n_ac / max(done, 1)}


REFERENCE_BUDGET_S = 150.0  # the whole --impl reference run (warm-up included) stays within a few minutes


def run_reference(args, rank, world):
    """Reference arm: the reference's CPU implementation of the path on the box's host cores. The reference itself
    cannot be built in this image, so this is the oracle port with all host threads. W untimed warm-up
    steps, then exactly K timed outer steps; each step is a bounded sample of config 2 (same case, coarser spacing),
    the spacing chosen so that K + W steps fit REFERENCE_BUDGET_S."""
    if rank != 0:
        return
    steps, warmup = max(args.steps, 1), max(args.warmup, 0)
    # one probe step at the finest sample resolution sizes the run: cost per outer step scales with the particle count
    dp = args.ref_dp
    probe = cpu_baseline_run(dp, 1, 1e9)
    while probe["_elapsed"] * (steps + warmup) > REFERENCE_BUDGET_S and dp < 0.05:
        probe["_elapsed"] /= 8.0
        dp *= 2.0
    r = cpu_baseline_run(dp, steps, 2.0 * REFERENCE_BUDGET_S, warmup=warmup)
    line = {
        "metric": "particle-steps/sec (3D WCSPH dam break)", "value": r["value"], "unit": "particle-steps/s",
        "n_gpus": args.gpus, "steps": r["_steps"], "warmup": warmup, "ms_per_step": r["_ms_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "impl": "reference",
        "config": {"workload": "3-D dam break WCSPH (tests_sycl dambreak geometry), AcousticRiemann + wall, Wendland C2 tabulated; "
                               f"bounded sample of config 2 (dp=0.00625) at dp={dp}: {r['_n_fluid']} fluid + {r['_n_wall']} wall particles",
                   "n_fluid_global": r["_n_fluid"], "n_wall": r["_n_wall"], "acoustic_steps_per_outer": r["_acoustic_per_outer"],
                   "parallelism": f"{r['cores']} host threads (OpenMP), rank 0 only"},
        "cpu_baseline": {k: v for k, v in r.items() if not k.startswith("_")},
        "e2e": {"value": r["value"], "unit": "particle-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def time_kernel(fn, iters, torch):
    """average launch duration of fn() over `iters` launches, CUDA events on the launching (current) stream."""
    fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters  # ms


def run_ours(args, rank, world, local_rank):
    import ctypes as C

    import torch
    import torch.distributed as dist
    from sphinxsys_b200 import host
    from sphinxsys_b200.host import DamBreakCK

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — libsphb200 has no CPU path (use --impl reference for the CPU oracle)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    # The C++ host layer builds the case itself (lattice generator + shape normals, include/sphinxsys_ck/dambreak_case.h)
    # and runs the case-file loop; Python only times it. Weak scaling: the SAME dam break refined
    # so that the fluid holds world x 4,096,000 particles (dp = 0.00625 / world^(1/3)), split into x-slabs of equal
    # particle count, one per GPU, with NCCL halo exchange of contiguous cell-plane ranges.
    dp = args.dp / (world ** (1.0 / 3.0)) if world > 1 else args.dp
    uid = None
    if world > 1:
        from sphinxsys_b200.host import comm_unique_id
        buf = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            buf.copy_(torch.frombuffer(bytearray(comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(buf, src=0)
        uid = bytes(buf.cpu().numpy().tobytes())
    solver = DamBreakCK(None, dim=3, dp=dp, device_index=local_rank, fused_time_step=True, sort_interval=100, generate=True,
                        rank=rank, nranks=world, unique_id=uid, serial_exchange=args.serial_exchange)
    solver.initialize()
    n_own = solver.own_range()[1]
    n_wall = solver.n_wall
    if world > 1:
        tn = torch.tensor([float(n_own)], dtype=torch.float64, device="cuda")
        dist.all_reduce(tn, op=dist.ReduceOp.SUM)
        n_fluid = int(tn.item())          # global fluid particles
    else:
        n_fluid = solver.n_fluid

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    solver.run_outer(args.warmup)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = solver.launches
    ac0 = solver.acoustic_steps
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    solver.run_outer(args.steps)  # K outer steps inside the C++ host loop (library work runs on the default stream)
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if rank == 0 else None
    n_ac = solver.acoustic_steps - ac0
    launches = solver.launches - launches0
    n_own = solver.own_range()[1]  # own particles of this rank after the timed steps (migration, re-cuts)
    t = torch.tensor([ms, float(n_ac)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)  # every rank takes the same sub-steps (global dt); time = slowest rank
        ms = float(t[0])
    total_particle_steps = n_fluid * float(n_ac)  # n_fluid is the GLOBAL particle count
    value = total_particle_steps / (ms * 1e-3)

    # ---- roofline of the dominant kernel (2nd-half fused launch), timed live on the same state ----
    peak, peak_kind = measured_peak()
    dt = solver.last_acoustic_dt
    ms_a2 = time_kernel(lambda: solver.exec("acoustic2", dt * 1e-3), 20, torch)
    ms_a1 = time_kernel(lambda: solver.exec("acoustic1", dt * 1e-3), 20, torch)
    ms_sum = time_kernel(lambda: solver.exec("density_summation"), 10, torch)
    ms_cl = time_kernel(lambda: solver.exec("rebuild"), 10, torch)  # cell list + storage reorder (+ migration/ghosts if decomposed)
    ms_rel = time_kernel(lambda: solver.exec("relations"), 5, torch)
    solver.exec("acoustic_dt_unprime")
    traffic = None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        traffic = tj.get("k_a2", {}).get(str(n_fluid))
    except Exception:
        pass
    ach_a2 = A_STEP_2ND * n_own / (ms_a2 * 1e-3) / 1e9  # per GPU
    roofline = {"bound": "hbm", "kernel": "k_a2 (AcousticStep2ndHalf: initialize+inner+wall+update+dt-max, one launch)",
                "achieved": ach_a2, "peak": peak, "peak_kind": peak_kind, "unit": "GB/s", "frac": ach_a2 / peak,
                "traffic": traffic, "algorithmic_bytes_per_launch": A_STEP_2ND * n_own, "launch_ms": ms_a2,
                "note": "pair arithmetic (~85 neighbours x ~50 instr, one 32-byte gather each) runs at 84 % of the L1 data-pipe and 72 % of the "
                        "issue peak (profiles/r01_v8_ncu_full.txt): bound by L1 bank wavefronts and FP32 issue, not by HBM; see DESIGN.md §4",
                "other_kernels_ms": {"acoustic_1st_half(init+interact)": ms_a1, "density_summation": ms_sum,
                                     "cell_list_build+reorder": ms_cl, "relation_build(inner+contact)": ms_rel}}

    # ---- e2e: the same step driven from HOST buffers (pinned), H2D of the evolving state + D2H of the result ----
    in_names = ["Position", "VolumetricMeasure", "Velocity", "Mass", "ForcePrior", "Compression", "CompressionRate",
                "VolumetricMeasureRef", "PreviousGravityForceCK"]
    out_names = ["Position", "Velocity", "Density"]

    decomposed = world > 1
    n_own = solver.own_range()[1]  # the own set as it is NOW (migration and re-cuts during the timed steps changed it)

    def pinned_like(name):
        # single GPU: the reference's packed layout in reference particle order; decomposed: this rank's own slots raw
        w = (4 if decomposed else 3) if name in host.VEC_NAMES else 1
        rows = n_own + n_own // 4 if decomposed else n_fluid
        tns = torch.empty((rows, w) if w > 1 else (rows,), dtype=torch.float32).pin_memory()
        return tns, tns.numpy()

    host_in = {nm: pinned_like(nm) for nm in in_names}
    host_out = {nm: pinned_like(nm) for nm in out_names}

    def fetch_inputs():
        for nm in in_names:
            if decomposed:
                solver.download_own_into(nm, host_in[nm][1])
            else:
                solver.download(nm, out=host_in[nm][1])

    fetch_inputs()
    per = lambda nm, vecw: (vecw if nm in host.VEC_NAMES else 1) * 4
    h2d = sum(per(nm, 4 if decomposed else 3) * (n_own if decomposed else n_fluid) for nm in in_names)
    d2h = sum(per(nm, 4 if decomposed else 3) * (n_own if decomposed else n_fluid) for nm in out_names)
    if decomposed:
        th = torch.tensor([float(h2d), float(d2h)], dtype=torch.float64, device="cuda")
        dist.all_reduce(th, op=dist.ReduceOp.SUM)
        h2d, d2h = int(th[0]), int(th[1])

    def e2e_step():
        if decomposed:
            for nm in in_names:
                solver.upload_own(nm, host_in[nm][1])   # own slots, storage order, from pinned memory
        else:
            for nm in in_names:
                solver.upload(nm, host_in[nm][1])       # DiscreteVariable::synchronizeToDevice (reference order)
        n = solver.step_outer()                         # configuration update (cell list, relations) FIRST, then the dynamics
        if decomposed:
            for nm in out_names:
                solver.download_own_into(nm, host_out[nm][1])
            # the configuration update ran at the START of the step on the uploaded state, so the rank still owns exactly
            # the particles of its host buffers (nothing migrated since): the same buffers are the next step's inputs
            assert solver.own_range()[1] == n_own, "own set changed during an e2e step"
        else:
            for nm in out_names:
                solver.download(nm, out=host_out[nm][1])  # DiscreteVariable::synchronizeWithDevice
        return n

    k_e2e = max(3, min(args.steps, 10))
    # the state arrives from the host every step: build the cell list and the relations for it at the START of the step
    # (DamBreakCK::ConfigurationUpdate::BeforeDynamics) instead of at its end — the same launches per step as `value`
    solver.exec("configuration_before_dynamics", 1.0)
    if decomposed:
        e2e_step()
        barrier()
        t0 = time.perf_counter()
        n_ac_e2e = 0
        for _ in range(k_e2e):
            n_ac_e2e += e2e_step()
        barrier()
        sec = time.perf_counter() - t0
        e2e_note = ("per step: H2D of all evolving variables of the rank's own particles from pinned host memory, migration + "
                    "cell-list + relation rebuild for the uploaded state, the dynamics of one outer step, D2H of Position/Velocity/Density")
    else:
        # single GPU: the same host buffers go through the HostTransferPipeline of the host layer — H2D of step s+1 and
        # D2H of step s-1 run on a side stream while step s computes; every step's copies are inside the timed region
        solver.pipeline_create(in_names, out_names)
        ins = [host_in[nm][1] for nm in in_names]
        outs = [host_out[nm][1] for nm in out_names]

        def run_pipelined(k):
            n_ac = 0
            solver.pipeline_stage_uploads(ins)
            for s_ in range(k):
                solver.pipeline_commit_uploads()
                if s_ + 1 < k:
                    solver.pipeline_stage_uploads(ins)   # next step's inputs: overlaps this step's dynamics
                n_ac += solver.step_outer()              # configuration update first (state came from the host), then dynamics
                solver.pipeline_stage_downloads(outs)    # overlaps the next step's dynamics
            solver.pipeline_synchronize()
            return n_ac

        run_pipelined(2)
        barrier()
        t0 = time.perf_counter()
        n_ac_e2e = run_pipelined(k_e2e)
        barrier()
        sec = time.perf_counter() - t0
        pb_in, pb_out = solver.pipeline_bytes()
        assert (pb_in, pb_out) == (h2d, d2h), (pb_in, h2d, pb_out, d2h)
        # the synchronous spelling (DiscreteVariable::synchronizeToDevice / synchronizeWithDevice per variable) for comparison
        e2e_step()
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        n_sync = sum(e2e_step() for _ in range(3))
        torch.cuda.synchronize()
        sec_sync = time.perf_counter() - t1
        e2e_note = ("per step: H2D of all evolving variables from pinned host memory (reference particle order), cell-list + "
                    "relation rebuild, one outer step, D2H of Position/Velocity/Density; copies run on a side stream and overlap "
                    "the neighbouring steps' dynamics (HostTransferPipeline); synchronous per-variable spelling: "
                    f"{n_fluid * float(n_sync) / sec_sync:.4g} particle-steps/s")
    tt = torch.tensor([sec, float(n_ac_e2e)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        sec = float(tt[0])
    tot = n_fluid * float(n_ac_e2e)
    e2e = {"value": tot / sec, "unit": "particle-steps/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
           "steps": k_e2e, "note": e2e_note}

    # ---- CPU baseline (rank 0, N=1 only): the oracle on the host cores, bounded sample ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        r = cpu_baseline_run(args.ref_dp, 3, 25.0)
        cpu = {k: v for k, v in r.items() if not k.startswith("_")}

    if rank == 0:
        n_ac_per_outer = n_ac / max(args.steps, 1)
        alg_bytes = A_STEP_1ST + A_STEP_2ND + A_OUTER / max(n_ac_per_outer, 1e-9) + A_SORT / (100.0 * max(n_ac_per_outer, 1e-9))
        line = {
            "metric": "particle-steps/sec (3D WCSPH dam break)", "value": value, "unit": "particle-steps/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / max(args.steps, 1),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"3-D dam break WCSPH (tests_sycl dambreak geometry), dp={dp:.6g}: {n_fluid} fluid + {n_wall} wall "
                                   f"particles, AcousticRiemann + wall, Wendland C2 tabulated, sort every 100 outer steps",
                       "n_fluid_global": n_fluid, "n_fluid_per_gpu": n_fluid // world, "n_wall": n_wall,
                       "acoustic_steps_per_outer": n_ac_per_outer,
                       "l2_policy": "working set (~2 GB per GPU incl. neighbour lists) larger than L2, no flush",
                       "parallelism": "1 GPU" if world == 1 else
                       f"{world} x-slabs of equal particle count on the global mesh, NCCL halo exchange of contiguous cell-plane ranges, "
                       f"full wall copy per rank"},
            "roofline": roofline,
            "whole_step_hbm_frac": value / world * alg_bytes / 1e9 / peak,  # per GPU
            "algorithmic_bytes_per_particle_step": alg_bytes,
            # also the un-amortised kernel-only rate (the two half-step launches of an acoustic step alone,
            # rank 0's launch times, every rank working on its own shard at once) and advection steps per second
            "kernel_only_particle_steps_per_s": n_fluid / ((ms_a1 + ms_a2) * 1e-3),
            "outer_steps_per_s": args.steps / (ms * 1e-3),
            "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--dp", type=float, default=0.00625, help="particle spacing; 0.00625 = config 2 (4,096,000 fluid)")
    ap.add_argument("--ref-dp", type=float, default=0.0125, help="resolution of the bounded CPU sample (512,000 fluid)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--serial-exchange", action="store_true", help="N > 1: plane exchange in line with the dynamics (no overlap)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        if args.warmup < 3:
            args.warmup = 3
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
