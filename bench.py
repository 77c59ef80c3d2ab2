#!/usr/bin/env python
"""bench.py — particle-steps/s of the 3-D WCSPH dam break (BASELINE.json metric) on N B200s of one node.

One "step" = one advection (outer) step of the case-file loop (dambreak.cpp:188-222): density summation +
regularisation, advection set-up, advection-dt reduction, the ~5 acoustic sub-steps it contains (acoustic-dt
reduction + 1st half + 2nd half each), position update, particle sort at its natural cadence (every 100 outer
steps), cell-linked-list and relation rebuild.  value = N_fluid x (acoustic sub-steps executed) / seconds, i.e.
every per-advection cost is amortised into the particle-step rate.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--dp 0.00625] [--impl ours|reference] [--no-extras]

Prints ONE JSON line (rank 0).  `--impl reference` times the CPU oracle (the reference cannot be built in this
image) on the box's host cores: the SAME case at the SAME spacing (config 2, dp = 0.00625) at every N, all host cores
(the thread count comes from the CPU affinity mask, not from OMP_NUM_THREADS, which torchrun sets to 1), bounded in the
number of advection steps only.

Sub-records of the line (all measured in the run): `parity` (N > 1: particle count, energy and an id-keyed 64-bit digest of
Position/Velocity after the warm-up, next to the same three from a 1-GPU run of the same spacing on rank 0's GPU),
`developed` (the same K steps timed again after --developed-steps more advection steps: the flow has left the lattice),
`complete_case` (correction variants + free-surface indication + probes: the case file as the reference runs it),
`config3` (N = 8: ~16 M fluid particles per GPU, dp = 0.002), `config4` (periodic Taylor-Green ring, n_side^3 per GPU),
`config5` (N = 1: neighbour-search chain at 16.7 M and 268 M random particles).
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


def host_threads():
    """Threads the CPU arm may use: the affinity mask of this process (torchrun exports OMP_NUM_THREADS=1, which would
    otherwise turn the all-cores baseline into a one-thread run)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except (AttributeError, OSError):
        return max(1, os.cpu_count() or 1)


def cpu_baseline_run(dp, max_outer, budget_s, threads=0, warmup=0, fit_s=None):
    """The oracle (CPU restatement of the reference's CK par_host data flow) on the host cores: bounded sample.
    `warmup` outer steps run untimed first; then up to `max_outer` timed outer steps (fewer only if budget_s runs out).
    fit_s: shrink warm-up and step count (never the case) so that set-up + warm-up + timed steps fit that many seconds;
    the first warm-up step is the probe."""
    from oracle import oracle as orc
    from sphinxsys_b200 import cases
    threads = threads or host_threads()
    os.environ["OMP_NUM_THREADS"] = str(threads)  # before the OpenMP runtime starts (first oracle call)
    t_setup = time.perf_counter()
    case = cases.dam_break(dim=3, dp=dp)
    sim = orc.OracleSim(case, f64=False, threads=threads)
    sim.exec("prepare_ck")
    t_setup = time.perf_counter() - t_setup
    warm_done = 0
    if fit_s is not None:
        t0 = time.perf_counter()
        sim.exec("run_ck", 1e9, 1, 1e9, 100)
        per_step = max(time.perf_counter() - t0, 1e-3)
        warm_done = 1
        fit = max(2, int((fit_s - t_setup - per_step) / per_step))
        if max_outer + max(warmup - 1, 0) > fit:
            warmup = 1 + min(max(warmup - 1, 0), fit // 5)
            max_outer = max(1, fit - (warmup - 1))
    for _ in range(max(warmup - warm_done, 0)):
        sim.exec("run_ck", 1e9, 1, 1e9, 100)
    warmup = max(warmup, warm_done)
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
    return {"value": case.n_fluid * n_ac / elapsed, "unit": "particle-steps/s", "cores": int(orc.lib().orc_max_threads()),
            "kind": "port", "ns_per_pair_interaction": 1e9 * elapsed / max(2.0 * pairs * n_ac, 1.0),
            "sample": f"3-D dam break dp={dp} ({case.n_fluid} fluid + {case.n_wall} wall), {done} outer / {n_ac} acoustic steps "
                      f"in {elapsed:.1f} s after {warmup} untimed (set-up {t_setup:.1f} s), oracle fp32 + OpenMP (restatement of the "
                      f"reference CK par_host path, not the TBB build)",
            "_elapsed": elapsed, "_steps": done, "_warmup": warmup, "_ms_per_step": 1e3 * elapsed / max(done, 1), "_n_fluid": case.n_fluid,
            "_n_wall": case.n_wall, "_acoustic_per_outer": n_ac / max(done, 1)}


REFERENCE_BUDGET_S = 150.0  # the whole --impl reference run (set-up and warm-up included) stays within a few minutes


def run_reference(args, rank, world):
    """Reference arm: the reference's CPU implementation of the path on the box's host cores. The reference itself
    cannot be built in this image, so this is the oracle port with all host threads, on config 2 at ITS
    OWN size (dp = 0.00625: 4,096,000 fluid + 3,034,688 wall) whatever N is — particle-steps/s of the CPU path does not
    depend on how many GPUs the other arm uses. W untimed warm-up steps, then K timed advection steps; the sample is
    bounded in the NUMBER of steps only (the first warm-up step sizes it so that the run fits REFERENCE_BUDGET_S)."""
    if rank != 0:
        return
    steps, warmup = max(args.steps, 1), max(args.warmup, 0)
    dp = args.ref_dp
    r = cpu_baseline_run(dp, steps, 2.0 * REFERENCE_BUDGET_S, threads=host_threads(), warmup=warmup, fit_s=REFERENCE_BUDGET_S)
    line = {
        "metric": "particle-steps/sec (3D WCSPH dam break)", "value": r["value"], "unit": "particle-steps/s",
        "n_gpus": args.gpus, "steps": r["_steps"], "warmup": r["_warmup"], "ms_per_step": r["_ms_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "impl": "reference",
        "config": {"workload": f"3-D dam break WCSPH (tests_sycl dambreak geometry), dp={dp:.6g}: {r['_n_fluid']} fluid + {r['_n_wall']} wall "
                               "particles, AcousticRiemann + wall, Wendland C2 tabulated (config 2 at its own size at every N; the sample "
                               f"is bounded in advection steps: {r['_steps']} timed of the {args.steps} asked for)",
                   "n_fluid_global": r["_n_fluid"], "n_wall": r["_n_wall"], "acoustic_steps_per_outer": r["_acoustic_per_outer"],
                   "parallelism": f"{r['cores']} host threads (OpenMP, from the affinity mask), rank 0 only"},
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


def kernel_counters():
    """ncu counters of the interaction kernels per fluid particle (profiles/kernel_counters.json, written from the committed
    ncu --set full captures): DRAM bytes, warp instructions and L1 data-pipe wavefronts of one launch divided by the particles
    it processed. They turn the launch time measured live into the second roofline (issue slots, L1 data pipe)."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "kernel_counters.json")))
    except Exception:
        return {}


def second_roofline(name, n_own, launch_ms, sm_mhz, counters):
    """Fractions of the two SM-side ceilings that bound the pair-interaction kernels: warp instructions per
    launch over the issue slots of the launch (148 SMs x 4 schedulers x cycles) and L1 data-pipe wavefronts per SM over its
    cycles (one wavefront per cycle and SM). Counters per particle come from the committed ncu capture of this kernel;
    the launch time and the SM clock are measured in this run."""
    c = counters.get(name)
    if not c or not sm_mhz or launch_ms <= 0:
        return None
    cycles = launch_ms * 1e-3 * sm_mhz * 1e6
    inst = c["warp_instructions_per_particle"] * n_own
    wave = c["l1_wavefronts_per_particle"] * n_own
    return {"bound": "issue slots / L1 data pipe", "issue_frac": inst / (cycles * 148 * 4), "l1_data_pipe_frac": wave / (cycles * 148),
            # the launch time each ceiling alone would allow (ms): what a perfect memory system / a perfect scheduler would leave
            "issue_floor_ms": 1e3 * inst / (148 * 4 * sm_mhz * 1e6), "l1_data_pipe_floor_ms": 1e3 * wave / (148 * sm_mhz * 1e6),
            "warp_instructions_per_launch": inst, "l1_wavefronts_per_launch": wave, "sm_mhz": sm_mhz, "source": c.get("source")}


def time_stream_kernels(solver, torch, n_own, peak, peak_kind, sm_mhz, counters):
    """Per-kernel launch times on the CURRENT state (CUDA events on the launching stream) and the roofline records."""
    dt = solver.last_acoustic_dt
    ms_a2 = time_kernel(lambda: solver.exec("acoustic2", dt * 1e-3), 20, torch)
    ms_a1 = time_kernel(lambda: solver.exec("acoustic1", dt * 1e-3), 20, torch)
    ms_sum = time_kernel(lambda: solver.exec("density_summation"), 10, torch)
    ms_cl = time_kernel(lambda: solver.exec("rebuild"), 10, torch)  # cell list + storage reorder (+ migration/ghosts if decomposed)
    ms_rel = time_kernel(lambda: solver.exec("relations"), 5, torch)
    solver.exec("acoustic_dt_unprime")
    traffic = None
    c2 = counters.get("k_a2")
    if c2:
        traffic = c2["dram_bytes_per_particle"] * n_own
    ach_a2 = A_STEP_2ND * n_own / (ms_a2 * 1e-3) / 1e9  # per GPU
    roofline = {"bound": "hbm", "kernel": "k_a2 (AcousticStep2ndHalf: initialize+inner+wall+update+dt-max, one launch)",
                "achieved": ach_a2, "peak": peak, "peak_kind": peak_kind, "unit": "GB/s", "frac": ach_a2 / peak,
                "traffic": traffic,
                "traffic_source": (c2 or {}).get("source", None),
                "algorithmic_bytes_per_launch": A_STEP_2ND * n_own, "launch_ms": ms_a2,
                "second": second_roofline("k_a2", n_own, ms_a2, sm_mhz, counters),
                "note": "stored-list gather design: ~85 neighbour rows per particle, one 32-byte per-lane gather and ~50 instructions each; "
                        "bound by the L1 data pipe and the issue rate (`second`), not by HBM; see DESIGN.md §4",
                "other_kernels_ms": {"acoustic_1st_half(init+interact)": ms_a1, "density_summation": ms_sum,
                                     "cell_list_build+reorder": ms_cl, "relation_build(inner+contact)": ms_rel},
                "other_kernels_second": {"k_a1_interact": second_roofline("k_a1_interact", n_own, ms_a1, sm_mhz, counters)}}
    return roofline, (ms_a1, ms_a2)


def bind_near_gpu(torch, index):
    """N > 1: run this rank (and allocate its pinned host buffers: first touch) on the NUMA node its GPU hangs off, so that
    the host <-> device copies of the e2e path do not cross the socket interconnect. Returns the node or None."""
    try:
        p = torch.cuda.get_device_properties(index)
        path = f"/sys/bus/pci/devices/{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0/numa_node"
        node = int(open(path).read().strip())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= set(os.sched_getaffinity(0))
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return node
    except Exception:
        return None


def run_ours(args, rank, world, local_rank):
    import ctypes as C

    import torch
    import torch.distributed as dist
    from sphinxsys_b200 import host
    from sphinxsys_b200.host import DamBreakCK

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — libsphb200 has no CPU path (use --impl reference for the CPU oracle)")
    torch.cuda.set_device(local_rank)
    numa_node = bind_near_gpu(torch, local_rank) if world > 1 else None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def new_unique_id():
        if world == 1:
            return None
        buf = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            buf.copy_(torch.frombuffer(bytearray(host.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(buf, src=0)
        return bytes(buf.cpu().numpy().tobytes())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(*vals):
        t = torch.tensor([float(v) for v in vals], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(x) for x in t]

    def sum_over_ranks(*vals):
        t = torch.tensor([float(v) for v in vals], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return [float(x) for x in t]

    def timed_outer(solver, k):
        """K advection steps inside the C++ host loop, CUDA events on the launching stream, barrier + synchronize on both
        sides, max over ranks. Returns (ms, acoustic steps, launches)."""
        l0, a0 = solver.launches, solver.acoustic_steps
        timed_outer.allocations_before = solver.device_allocations
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        ev0.record()
        solver.run_outer(k)
        ev1.record()
        barrier()
        ms, = max_over_ranks(ev0.elapsed_time(ev1))
        timed_outer.allocations = solver.device_allocations - timed_outer.allocations_before
        return ms, solver.acoustic_steps - a0, solver.launches - l0

    # The C++ host layer builds the case itself (lattice generator + shape normals, include/sphinxsys_ck/dambreak_case.h)
    # and runs the case-file loop; Python only times it. Weak scaling: the SAME dam break refined
    # so that the fluid holds world x 4,096,000 particles (dp = 0.00625 / world^(1/3)), split into x-slabs of equal
    # particle count, one per GPU, with NCCL halo exchange of contiguous cell-plane ranges.
    dp = args.dp / (world ** (1.0 / 3.0)) if world > 1 else args.dp
    # the particle sort (decomposed: the re-cut of the slabs) runs every 100 advection steps in the case file; a timed window
    # shorter than that gets exactly ONE inside it (cadence = K), i.e. the sort weighs 100/K times its natural share
    cadence = min(100, max(args.steps, 2))
    solver = DamBreakCK(None, dim=3, dp=dp, device_index=local_rank, fused_time_step=True, sort_interval=cadence, generate=True,
                        rank=rank, nranks=world, unique_id=new_unique_id(), serial_exchange=args.serial_exchange, recut_interval=cadence)
    solver.initialize()
    n_wall = solver.n_wall
    n_wall_global = int(solver.exec("wall_global_particles"))
    n_fluid = int(sum_over_ranks(solver.own_range()[1])[0]) if world > 1 else solver.n_fluid

    if world == 1:
        # the first ParticleSortCK allocates its key / permutation buffers: outside the timed window (the state is unchanged:
        # a sort only renumbers)
        solver.exec("sort")
        solver.exec("rebuild")
        solver.exec("relations")
    solver.run_outer(args.warmup)
    barrier()

    # ---- parity evidence carried by the multi-GPU numbers: the state after the warm-up against ONE GPU ----
    parity = None
    if world > 1 and not args.no_parity:
        cnt, dig = solver.state_digest()
        gathered = [None] * world
        dist.all_gather_object(gathered, (cnt, dig))
        energy = solver.energy()        # all-reduced inside the host layer
        ac_dec = solver.acoustic_steps
        one = None
        if rank == 0:
            try:
                single = DamBreakCK(None, dim=3, dp=dp, device_index=local_rank, fused_time_step=True, sort_interval=cadence, generate=True)
                single.initialize()
                single.run_outer(args.warmup)
                c1, d1 = single.state_digest()
                one = {"particles": c1, "energy": single.energy(), "digest": f"{d1:016x}", "acoustic_steps": single.acoustic_steps}
                single.close()
                del single
                torch.cuda.empty_cache()
            except Exception as e:  # e.g. the global case does not fit one GPU
                one = {"error": str(e)[:300]}
        barrier()
        if rank == 0:
            tot = sum(c for c, _ in gathered)
            dsum = sum(d for _, d in gathered) % (1 << 64)
            parity = {"after_outer_steps": args.warmup, "decomposed": {"particles": tot, "energy": energy, "digest": f"{dsum:016x}",
                                                                        "acoustic_steps": ac_dec, "own_per_rank": [c for c, _ in gathered]},
                      "single_gpu": one,
                      "bit_identical": bool(one and one.get("digest") == f"{dsum:016x}" and one.get("particles") == tot),
                      "digest": "sum mod 2^64 over particles of mix64(global id, Position bits, Velocity bits): order- and partition-independent"}

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms, n_ac, launches = timed_outer(solver, args.steps)
    allocations_in_window = timed_outer.allocations
    clocks = sampler.stop() if rank == 0 else None
    sorts_in_window = sum(1 for it in range(args.warmup + 1, args.warmup + args.steps + 1) if it % cadence == 0 and it != 1)
    n_own = solver.own_range()[1]  # own particles of this rank after the timed steps (migration, re-cuts)
    total_particle_steps = n_fluid * float(n_ac)  # n_fluid is the GLOBAL particle count
    value = total_particle_steps / (ms * 1e-3)

    # ---- roofline of the dominant kernel (2nd-half fused launch), timed live on the same state ----
    peak, peak_kind = measured_peak()
    counters = kernel_counters()
    sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
    roofline, (ms_a1, ms_a2) = time_stream_kernels(solver, torch, n_own, peak, peak_kind, sm_mhz, counters)
    # one ParticleSortCK (keys + radix sort + renumbering) and the configuration update that follows it, timed alone
    sort_ms = None
    if world == 1:
        sort_ms = time_kernel(lambda: solver.exec("sort"), 3, torch)
        solver.exec("rebuild")
        solver.exec("relations")

    # (the developed-state leg runs on its own solver AFTER the e2e leg — developed_leg() below: the e2e leg takes the state the
    # timed window left, i.e. the kind of state `value` was measured on, at every N. It used to inherit the developed state,
    # which at the finer spacings of the N-GPU runs holds 2-3 acoustic steps per advection step instead of 5: an e2e rate in
    # particle-STEPS per second that could not be compared with the one at N = 1.)
    # ---- e2e: the same step driven from HOST buffers (pinned), H2D of the evolving state + D2H of the result ----
    in_names = ["Position", "VolumetricMeasure", "Velocity", "Mass", "ForcePrior", "Compression", "CompressionRate",
                "VolumetricMeasureRef", "PreviousGravityForceCK"]
    out_names = ["Position", "Velocity", "Density"]

    decomposed = world > 1
    n_own = solver.own_range()[1]  # the own set as it is NOW (migration and re-cuts during the timed steps changed it)

    def pinned_like(name):
        # the reference's packed layout (Vecd = 3 floats). Single GPU: reference particle order; decomposed: this rank's own
        # particles in slot order (room for a quarter more: the own set may grow)
        w = 3 if name in host.VEC_NAMES else 1
        rows = n_own + n_own // 4 if decomposed else n_fluid
        tns = torch.empty((rows, w) if w > 1 else (rows,), dtype=torch.float32).pin_memory()
        return tns, tns.numpy()

    host_in = {nm: pinned_like(nm) for nm in in_names}
    host_out = {nm: pinned_like(nm) for nm in out_names}

    def fetch_inputs():
        for nm in in_names:
            if decomposed:
                host_in[nm][1][:n_own] = solver.download_own(nm)
            else:
                solver.download(nm, out=host_in[nm][1])

    fetch_inputs()
    per = lambda nm: (3 if nm in host.VEC_NAMES else 1) * 4
    h2d = sum(per(nm) * (n_own if decomposed else n_fluid) for nm in in_names)
    d2h = sum(per(nm) * (n_own if decomposed else n_fluid) for nm in out_names)
    h2d_own, d2h_own = h2d, d2h
    if decomposed:
        h2d, d2h = (int(v) for v in sum_over_ranks(h2d, d2h))

    def e2e_step():
        # the synchronous spelling, single GPU only: DiscreteVariable::synchronizeToDevice / synchronizeWithDevice per variable
        for nm in in_names:
            solver.upload(nm, host_in[nm][1])
        n = solver.step_outer()                         # configuration update (cell list, relations) FIRST, then the dynamics
        for nm in out_names:
            solver.download(nm, out=host_out[nm][1])
        return n

    k_e2e = max(3, min(args.steps, 10))
    # the state arrives from the host every step: build the cell list and the relations for it at the START of the step
    # (DamBreakCK::ConfigurationUpdate::BeforeDynamics) instead of at its end — the same launches per step as `value`
    solver.exec("configuration_before_dynamics", 1.0)
    solver.exec("set_sort_interval", 0)  # the state is re-uploaded every step: no renumbering / re-cut in between
    # The host buffers go through the HostTransferPipeline of the host layer — H2D of step s+1 and D2H of step s-1 run on a
    # side stream while step s computes; every step's copies are inside the timed region. Single GPU: reference particle
    # order and packed layout; decomposed: this rank's own slots, raw (HostTransferPipeline::setRawOwnSlots).
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
            if decomposed:
                assert solver.own_range()[1] == n_own, "own set changed during an e2e step"
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
    assert (pb_in, pb_out) == (h2d_own, d2h_own), (pb_in, h2d_own, pb_out, d2h_own)
    # the ceiling the host side sets: the same copies with NO dynamics in between, all ranks at once (PCIe + host memory)
    def transfer_only(k):
        for _ in range(k):
            solver.pipeline_stage_uploads(ins)
            solver.pipeline_commit_uploads()
            solver.pipeline_stage_downloads(outs)
        solver.pipeline_synchronize()
        torch.cuda.synchronize()
    transfer_only(1)
    barrier()
    t2 = time.perf_counter()
    transfer_only(4)
    barrier()
    sec_copy, = max_over_ranks((time.perf_counter() - t2) / 4)
    sec, = max_over_ranks(sec)
    sync_note = ""
    if not decomposed:
        # the synchronous spelling (one blocking copy per variable) for comparison
        e2e_step()
        barrier()
        t1 = time.perf_counter()
        n_sync = sum(e2e_step() for _ in range(3))
        barrier()
        sec_sync = time.perf_counter() - t1
        sync_note = f"; synchronous per-variable spelling: {n_fluid * float(n_sync) / sec_sync:.4g} particle-steps/s"
    e2e_note = ("per step: H2D of all evolving variables from pinned host memory (" +
                ("this rank's own particles in slot order" if decomposed else "reference particle order") +
                ", Vecd packed as 3 floats), cell-list + relation rebuild" + (" with migration and ghost-plane exchange" if decomposed else "") +
                ", one outer step, D2H of Position/Velocity/Density; copies run on a side stream and overlap the neighbouring steps' "
                "dynamics (HostTransferPipeline)" + sync_note +
                (f"; rank bound to NUMA node {numa_node} of its GPU" if numa_node is not None else ""))
    tot = n_fluid * float(n_ac_e2e)
    e2e = {"value": tot / sec, "unit": "particle-steps/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
           "steps": k_e2e, "ms_per_step": 1e3 * sec / k_e2e,
           # copies alone (no dynamics), slowest rank, all ranks copying at once: what host memory + PCIe allow per step
           "transfer_only_ms_per_step": 1e3 * sec_copy,
           "transfer_only_gb_per_s_per_gpu": (h2d_own + d2h_own) / sec_copy / 1e9,
           "note": e2e_note}

    solver.close()
    del solver, host_in, host_out, ins, outs
    torch.cuda.empty_cache()

    # ---- CPU baseline (rank 0, N=1 only): the oracle on the host cores, bounded sample of the SAME case and spacing ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        r = cpu_baseline_run(args.ref_dp, 2, 40.0, threads=host_threads(), warmup=1)
        cpu = {k: v for k, v in r.items() if not k.startswith("_")}

    # ---- the other BASELINE configs, each a small driver-visible record (failures are reported, not fatal) ----
    extras = {}
    developed = None
    if not args.no_extras and args.developed_steps > 0:
        developed = guarded_leg(lambda: developed_leg(args, dp, cadence, rank, world, local_rank, torch, new_unique_id, timed_outer,
                                                      sum_over_ranks, max_over_ranks, n_fluid, peak, peak_kind, sm_mhz, counters), rank)
        barrier()
    if not args.no_extras:
        extras["complete_case"] = guarded_leg(lambda: complete_case_leg(args, dp, rank, world, local_rank, torch, new_unique_id, timed_outer,
                                                                       sum_over_ranks), rank)
        barrier()
        if world == 8 or args.force_config3:
            extras["config3"] = guarded_leg(lambda: config3_leg(args, rank, world, local_rank, torch, dist, new_unique_id, timed_outer,
                                                               sum_over_ranks, max_over_ranks), rank)
            barrier()
        extras["config4"] = guarded_leg(lambda: config4_leg(args, rank, world, local_rank, torch, dist, new_unique_id, barrier,
                                                           max_over_ranks), rank)
        barrier()
        if world == 1:
            extras["config5"] = guarded_leg(lambda: config5_leg(args), rank)

    if rank == 0:
        n_ac_per_outer = n_ac / max(args.steps, 1)
        alg_bytes = A_STEP_1ST + A_STEP_2ND + A_OUTER / max(n_ac_per_outer, 1e-9) + A_SORT / (100.0 * max(n_ac_per_outer, 1e-9))
        line = {
            "metric": "particle-steps/sec (3D WCSPH dam break)", "value": value, "unit": "particle-steps/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / max(args.steps, 1),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"3-D dam break WCSPH (tests_sycl dambreak geometry), dp={dp:.6g}: {n_fluid} fluid + {n_wall} wall "
                                   f"particles, AcousticRiemann + wall, Wendland C2 tabulated; " +
                                   (f"slab re-cut every {cadence}" if world > 1 else f"ParticleSortCK every {cadence}") +
                                   f" advection steps ({sorts_in_window} inside the timed window; the case file's cadence is 100)",
                       "n_fluid_global": n_fluid, "n_fluid_per_gpu": n_fluid // world, "n_wall": n_wall,
                       "acoustic_steps_per_outer": n_ac_per_outer, "sorts_in_timed_window": sorts_in_window, "sort_ms": sort_ms,
                       "l2_policy": "working set (~2 GB per GPU incl. neighbour lists) larger than L2, no flush",
                       "parallelism": "1 GPU" if world == 1 else
                       f"{world} x-slabs of equal particle count on the global mesh; halo refresh: NCCL send/recv of contiguous cell-plane "
                       f"ranges; migration: peer-mailbox writes over NVLink (CUDA IPC); every rank stores its slab of the wall "
                       f"(n_wall = rank 0's share of {n_wall_global})"},
            "roofline": roofline,
            "whole_step_hbm_frac": value / world * alg_bytes / 1e9 / peak,  # per GPU
            "algorithmic_bytes_per_particle_step": alg_bytes,
            # also the un-amortised kernel-only rate (the two half-step launches of an acoustic step alone,
            # rank 0's launch times, every rank working on its own shard at once) and advection steps per second
            "kernel_only_particle_steps_per_s": n_fluid / ((ms_a1 + ms_a2) * 1e-3),
            "outer_steps_per_s": args.steps / (ms * 1e-3),
            "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
            "device_allocations_in_timed_window_rank0": int(allocations_in_window),
            "parity": parity, "developed": developed,
        }
        line.update(extras)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def guarded_leg(fn, rank):
    """A sub-record must never take the headline down: an exception becomes {"error": ...} (every rank runs the leg)."""
    try:
        return fn()
    except Exception as e:  # noqa: BLE001
        import traceback
        if rank == 0:
            traceback.print_exc(file=sys.stderr)
        return {"error": f"{type(e).__name__}: {str(e)[:300]}"}


def developed_leg(args, dp, cadence, rank, world, local_rank, torch, new_unique_id, timed_outer, sum_over_ranks, max_over_ranks, n_fluid,
                  peak, peak_kind, sm_mhz, counters):
    """the same K steps timed after the flow has left the initial lattice (--developed-steps more advection steps:
    free surface formed, fewer neighbours per particle, re-cuts done) — on a solver of its own, same case and spacing."""
    from sphinxsys_b200.host import DamBreakCK
    solver = DamBreakCK(None, dim=3, dp=dp, device_index=local_rank, fused_time_step=True, sort_interval=100, generate=True,
                        rank=rank, nranks=world, unique_id=new_unique_id(), serial_exchange=args.serial_exchange, recut_interval=100)
    solver.initialize()
    solver.run_outer(args.warmup + args.steps + args.developed_steps)
    solver.exec("set_sort_interval", cadence)
    sampler2 = ClockSampler(local_rank)
    if rank == 0:
        sampler2.start()
    ms_d, n_ac_d, launches_d = timed_outer(solver, args.steps)
    clocks_d = sampler2.stop() if rank == 0 else None
    n_own_d = solver.own_range()[1]
    roof_d, (ms_a1_d, ms_a2_d) = time_stream_kernels(solver, torch, n_own_d, peak, peak_kind, (clocks_d or {}).get("sm_mhz") or sm_mhz, counters)
    pairs_in, pairs_ct = sum_over_ranks(solver.exec("inner_pairs"), solver.exec("contact_pairs"))
    own_total, own_max = sum_over_ranks(n_own_d)[0], max_over_ranks(n_own_d)[0]
    developed = {"after_outer_steps": args.warmup + args.steps + args.developed_steps, "physical_time": solver.physical_time,
                 "value": n_fluid * float(n_ac_d) / (ms_d * 1e-3), "unit": "particle-steps/s", "ms_per_step": ms_d / max(args.steps, 1),
                 "ms_per_acoustic_step": ms_d / max(n_ac_d, 1),
                 "acoustic_steps_per_outer": n_ac_d / max(args.steps, 1), "gpu_launches": int(launches_d),
                 "inner_neighbours_per_particle": pairs_in / max(own_total, 1.0),
                 "wall_neighbours_per_particle": pairs_ct / max(own_total, 1.0),
                 "load_imbalance_max_over_mean": own_max / max(own_total / world, 1.0),
                 "k_a2_ms": ms_a2_d, "k_a1_ms": ms_a1_d, "roofline_frac_k_a2": roof_d["frac"], "second": roof_d["second"],
                 "other_kernels_ms": roof_d["other_kernels_ms"], "clocks": clocks_d}
    solver.close()
    del solver
    torch.cuda.empty_cache()
    return developed


def complete_case_leg(args, dp, rank, world, local_rank, torch, new_unique_id, timed_outer, sum_over_ranks):
    """The case file as the reference actually runs it (dambreak.cpp:117-134,192-193,223-224): LinearCorrectionMatrix + the
    Correction aliases of both half steps, FreeSurfaceIndicationComplexSpatialTemporalCK every advection step, the six
    wall-pressure probes interpolated and recorded every advection step — same particles as `value`."""
    from sphinxsys_b200.host import DamBreakCK
    k, w = max(3, min(args.steps, 20)), 3
    s = DamBreakCK(None, dim=3, dp=dp, device_index=local_rank, fused_time_step=True, sort_interval=100, generate=True, correction=True,
                   surface_indicator=True, observers=True, rank=rank, nranks=world, unique_id=new_unique_id(), recut_interval=100)
    s.initialize()
    n, = sum_over_ranks(s.own_range()[1])
    s.run_outer(w)
    ms, n_ac, launches = timed_outer(s, k)
    rec = {"dynamics": "AcousticStep1st/2ndHalfWithWallRiemannCorrectionCK, LinearCorrectionMatrixComplex, "
                       "FreeSurfaceIndicationComplexSpatialTemporalCK, 6 pressure probes per advection step",
           "n_fluid_global": int(n), "steps": k, "warmup": w, "ms_per_step": ms / k, "acoustic_steps_per_outer": n_ac / k,
           "value": n * n_ac / (ms * 1e-3), "unit": "particle-steps/s", "gpu_launches": int(launches), "energy": s.energy(),
           "probe_records": int(s.exec("probe_records"))}
    s.close()
    del s
    torch.cuda.empty_cache()
    return rec


def config3_leg(args, rank, world, local_rank, torch, dist, new_unique_id, timed_outer, sum_over_ranks, max_over_ranks):
    """BASELINE config 3 at its own size: the dam break at dp = 0.002 (125 M fluid + ~29 M wall) over the GPUs of the box,
    ~16 M fluid particles per GPU."""
    from sphinxsys_b200.host import DamBreakCK
    dp3 = args.config3_dp
    k, w = max(3, min(args.steps, 10)), 3
    t0 = time.perf_counter()
    s3 = DamBreakCK(None, dim=3, dp=dp3, device_index=local_rank, fused_time_step=True, sort_interval=100, generate=True,
                    rank=rank, nranks=world, unique_id=new_unique_id(), recut_interval=100)
    s3.initialize()
    setup_s = time.perf_counter() - t0
    n3, = sum_over_ranks(s3.own_range()[1])
    stored_max, = max_over_ranks(s3.own_range()[2])
    s3.run_outer(w)
    ms3, n_ac3, l3 = timed_outer(s3, k)
    cnt, dig = s3.state_digest()
    tot, = sum_over_ranks(cnt)
    e3 = s3.energy()
    mem = torch.cuda.max_memory_allocated()  # torch's own share only; the library allocates with cudaMalloc
    free_b, total_b = torch.cuda.mem_get_info()
    rec = {"dp": dp3, "n_fluid_global": int(n3), "n_wall_global": int(s3.exec("wall_global_particles")), "n_wall_stored_rank0": s3.n_wall,
           "rebuild_host_round_trips": int(s3.exec("rebuild_host_syncs")), "n_fluid_per_gpu": int(n3) // world, "stored_per_gpu_max": int(stored_max),
           "steps": k, "warmup": w, "ms_per_step": ms3 / k, "acoustic_steps_per_outer": n_ac3 / k,
           "value": n3 * n_ac3 / (ms3 * 1e-3), "unit": "particle-steps/s", "gpu_launches": int(l3), "setup_s": setup_s,
           "particles_after": int(tot), "energy": e3, "hbm_used_gb_rank0": (total_b - free_b) / 1e9, "_torch_peak": mem}
    rec.pop("_torch_peak")
    s3.close()
    del s3
    torch.cuda.empty_cache()
    return rec


def config4_leg(args, rank, world, local_rank, torch, dist, new_unique_id, barrier, max_over_ranks):
    """BASELINE config 4: periodic Taylor-Green vortex, n_side^3 particles PER GPU, the box replicated along x as a ring of
    `world` slabs (x through the slab exchange + seam shift, y / z by image particles); one GPU: the ring of one slab."""
    import dataclasses
    from sphinxsys_b200 import cases, host
    n_side = args.config4_side
    k, w = max(4, min(args.steps, 8)), 4
    t0 = time.perf_counter()
    case = cases.taylor_green(dim=3, n_side=n_side, jitter=0.05)
    n_local = case.n_fluid
    local_ids = None
    if world > 1:
        pos = case.fluid_pos.copy()
        pos[:, 0] += np.float32(rank)
        up = list(case.periodic_upper)
        up[0] = float(world)
        case = dataclasses.replace(case, fluid_pos=pos, DL=float(world), LL=float(world), periodic_upper=tuple(up))
        local_ids = np.arange(n_local, dtype=np.uint32) + np.uint32(rank * n_local)
    gpu = host.TaylorGreenCK(case, device_index=local_rank, ring=True, rank=rank, nranks=world, unique_id=new_unique_id(),
                             local_ids=local_ids, sort_interval=0)
    gpu.initialize()
    setup_s = time.perf_counter() - t0
    gpu.run_outer(w)
    gpu.synchronize()
    l0 = gpu.launches
    # one event per advection step: a fresh process / freshly returned device memory makes the first steps of a large case
    # several times slower than the steady state (seen: 815 -> 43 ms per step at 256^3), so the record carries every step
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(k + 1)]
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    barrier()
    evs[0].record()
    n_ac = 0
    allocs = [gpu.device_allocations]
    for s_ in range(k):
        n_ac += gpu.run_outer(1)
        evs[s_ + 1].record()
        allocs.append(gpu.device_allocations)
    barrier()
    clocks4 = sampler.stop() if rank == 0 else None
    per_step = [evs[i].elapsed_time(evs[i + 1]) for i in range(k)]
    ms4, = max_over_ranks(evs[0].elapsed_time(evs[k]))
    med4, = max_over_ranks(float(np.median(per_step)))
    dt4 = gpu.last_acoustic_dt
    parts = {"acoustic_2nd_half": time_kernel(lambda: gpu.exec("acoustic2", dt4 * 1e-3), 5, torch),
             "acoustic_1st_half": time_kernel(lambda: gpu.exec("acoustic1", dt4 * 1e-3), 5, torch),
             "density_summation": time_kernel(lambda: gpu.exec("density_summation"), 3, torch),
             "cell_list+migration+ghost_planes+images": time_kernel(lambda: gpu.exec("rebuild"), 3, torch),
             "relation_build": time_kernel(lambda: gpu.exec("relations"), 3, torch),
             "inner_stride": gpu.exec("inner_stride"), "inner_max_count": gpu.exec("inner_max_count")}
    rec = {"n_side": n_side, "kernels_ms_rank0": parts, "particles_per_gpu": n_local, "n_fluid_global": n_local * world, "steps": k, "warmup": w,
           "ms_per_step": ms4 / k, "ms_per_step_rank0": [round(v, 2) for v in per_step],
           "device_allocations_per_step_rank0": [b - a for a, b in zip(allocs[:-1], allocs[1:])], "acoustic_steps_per_outer": n_ac / k,
           "value": world * n_local * n_ac / (ms4 * 1e-3),
           # the same from the MEDIAN step (slowest rank): single steps of this leg come out 1.5-2.7x slow now and then
           "ms_per_step_median": med4, "value_median_step": world * n_local * (n_ac / k) / (med4 * 1e-3), "clocks": clocks4,
           "unit": "particle-steps/s", "gpu_launches": gpu.launches - l0, "images_rank0": gpu.ghost_particles,
           "plane_ghosts_rank0": int(gpu.exec("plane_ghost_particles")), "kinetic_energy": gpu.energy(), "setup_s": setup_s,
           "parallelism": f"ring of {world} slab(s) along x, NCCL" if world > 1 else "ring of one slab (device copies)"}
    gpu.close()
    del gpu
    torch.cuda.empty_cache()
    return rec


def config5_leg(args):
    """BASELINE config 5: neighbour-search chain (keys, radix sort, permutation, cell list + reorder, neighbour count) on
    random particles, through scripts/config5_bench.py in a child process (its own context; bounded by a timeout)."""
    out = os.path.join(ROOT, "gpurun_out", "config5_bench_leg.json")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    p = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "config5_bench.py"), "--sizes", args.config5_sizes, "--reps", "2",
                        "--out", out], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=240)
    if p.returncode != 0 and not os.path.exists(out):
        return {"error": (p.stderr or p.stdout)[-300:]}
    d = json.load(open(out))
    rows = [{"particles": r["particles"], "ms_total": r["ms_total"], "particles_per_s": r["particles_per_s"], "ms": r["ms"],
             "hbm_frac_algorithmic": r["hbm_frac_algorithmic"], "properties_ok": r["properties_ok"]} for r in d["rows"]]
    return {"algorithmic_bytes_per_particle": d["algorithmic_bytes_per_particle"], "rows": rows}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--dp", type=float, default=0.00625, help="particle spacing; 0.00625 = config 2 (4,096,000 fluid)")
    ap.add_argument("--ref-dp", type=float, default=0.00625, help="spacing of the CPU arm: config 2 at its own size (4,096,000 fluid)")
    ap.add_argument("--developed-steps", type=int, default=1200, help="advection steps run (untimed) before the developed-state leg")
    ap.add_argument("--no-extras", action="store_true", help="skip the developed / config3 / config4 / config5 sub-records")
    ap.add_argument("--config4-side", type=int, default=256, help="particles per side and GPU of the config-4 ring leg")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--serial-exchange", action="store_true", help="N > 1: plane exchange in line with the dynamics (no overlap)")
    ap.add_argument("--no-parity", action="store_true", help="N > 1: skip the 1-GPU comparison run of the parity sub-record")
    ap.add_argument("--config3-dp", type=float, default=0.002, help="spacing of the config-3 leg (0.002: 125 M fluid particles)")
    ap.add_argument("--force-config3", action="store_true", help="run the config-3 leg at any N (default: N = 8 only)")
    ap.add_argument("--config5-sizes", default="16,256", help="config-5 leg: millions (2^20) of random particles")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        if args.warmup < 3:
            args.warmup = 3
        try:
            run_ours(args, rank, world, local_rank)
        except SystemExit:
            raise
        except BaseException:
            # N > 1: a rank-local failure must not leave the peers waiting in a collective until some outer timeout
            import traceback
            traceback.print_exc()
            sys.stderr.flush()
            if world > 1:
                os._exit(1)
            raise


if __name__ == "__main__":
    main()
