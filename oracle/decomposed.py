"""Slab-decomposed CPU restatement of the N-GPU run, built on the single-domain oracle.

TEST INFRASTRUCTURE: imported only by tests/. The product package `sphinxsys_b200` never imports this module.

What it pins. The reference has no distributed path; the decomposed run is new (include/sphinxsys_ck/slab_decomposition.h,
DamBreakCK::stepOuter). Its contract is "every rank's own particles carry, bit for bit, the values the undecomposed run
gives them". This module restates the EXCHANGE PROTOCOL of that run on the CPU — who owns a particle, what travels at
the configuration update, which variables are refreshed on the ghost planes between the stages of a step — with the
oracle (oracle/sph_oracle.cpp) doing the arithmetic of each stage on `own + ghost` particles, so that the protocol can
be checked against the single-domain oracle without a GPU (tests/test_decomposed_oracle_cpu.py, world_size-2 gloo).

Protocol (one rank; same order as DamBreakCK::stepOuter / SlabDecomposition::rebuild):
  * ownership: rank r owns the cell planes [cut[r], cut[r+1]) of the GLOBAL mesh along x (the slowest cell axis);
  * configuration update: particles whose plane left the slab go to their new owner with every variable (they move
    less than one plane per advection step, so the owner is a neighbour rank: asserted); the first / last own plane
    is sent to the left / right neighbour as its ghost plane, every variable;
  * stored order `own | ghosts from the left | ghosts from the right`, each by ascending global id, so the order
    inside a cell — hence every neighbour row and every summation — is the single-domain one;
  * ghost refresh inside a step: VolumetricMeasure after AdvectionStepSetup, Pressure after the initialisation of
    the 1st half, Velocity after its update, LinearCorrectionMatrix after it is rebuilt (correction variants only),
    PositionDivergence between the two sweeps of the free-surface indication (when the case has it);
    nothing else (the viscous force, the kernel gradient integral and the transport-velocity correction of the
    Taylor-Green case run on what is already there). Stages also run on the ghosts here (the GPU runs them on the
    active range only); what they leave there is meaningless and never read — except where a ghost close to the cut
    would get the RIGHT value by accident (PositionDivergence), which is wiped before the refresh so that a missing
    refresh cannot hide;
  * observer probes: recorded by the rank that owns the probe's cell plane (it stores the whole neighbourhood);
  * time steps: raw reductions over the OWN particles, max over the ranks, then the CFL formula.
The GPU differs from this restatement in one place: on a ring it decides what left the box by cell PLANE and keeps a
shifted particle inside the plane it belongs to (sphb200_seam_shift), while the wrap below is by position as in the
reference; the two agree except for particles within rounding of a face (identical modulo L).
Periodic runs along x ("ring"): the box is a whole number of cell planes (aligned mesh, spacing >= cut-off), the
first and the last rank are neighbours, ghost planes that cross the seam travel with Position shifted by -/+ L
(the arithmetic of the reference's ghost list entry, domain_bounding.cpp:26,45); y / z periodicity stays local (the
oracle's image entries, made for own + ghost particles alike); bounding wraps x by POSITION first
(domain_bounding.h:98-108), ownership follows the plane of the wrapped position, a particle exactly on the upper
bound stays with the last rank.
"""
from __future__ import annotations

import dataclasses
import threading

import numpy as np

from . import oracle as orc

# every per-particle variable of the CK formulation without kernel correction (oracle ensureFluidState)
VARIABLES = [("Position", 3), ("Velocity", 3), ("Displacement", 3), ("Force", 3), ("ForcePrior", 3),
             ("PreviousGravityForceCK", 3), ("VolumetricMeasure", 1), ("VolumetricMeasureRef", 1), ("Mass", 1),
             ("Density", 1), ("Pressure", 1), ("Compression", 1), ("CompressionRate", 1), ("CompressionSummation", 1)]
# carried as well when the case has viscosity (ForcePriorCK keeps the previous viscous force to form the increment)
VISCOUS_VARIABLES = [("ViscousForce", 3), ("PreviousViscousForce", 3)]
# LinearCorrectionCK variants: B is rebuilt every advection step, but the viscous force of the NEXT step still reads it
CORRECTION_VARIABLES = [("LinearCorrectionMatrix", 9)]
# FreeSurfaceIndicationCK: PositionDivergence lives inside one call (refreshed on the ghost planes between its two sweeps);
# the indicator of the previous step is an evolving variable (surface_indication_ck.hpp:42-46) and travels with the particle
INDICATOR_UINTS = ["PreviousSurfaceIndicator", "Indicator"]
WIDTH = dict(VARIABLES + VISCOUS_VARIABLES + CORRECTION_VARIABLES + [("PositionDivergence", 1)])


# ------------------------------------------------------------------------------------------------------------------
# transports: route(parcels) delivers {destination rank: payload} and returns {source rank: payload}
# ------------------------------------------------------------------------------------------------------------------
class SerialComm:
    rank, size = 0, 1

    def route(self, parcels):
        return {0: parcels[0]} if 0 in parcels else {}

    def allreduce_max(self, v):
        return float(v)


class GlooComm:
    """torch.distributed (gloo) transport: one process per rank, as the GPU run has one process per GPU."""

    def __init__(self):
        import torch.distributed as dist
        self._dist = dist
        self.rank, self.size = dist.get_rank(), dist.get_world_size()

    def route(self, parcels):
        boxes = [None] * self.size
        self._dist.all_gather_object(boxes, parcels)
        return {src: box[self.rank] for src, box in enumerate(boxes) if self.rank in box}

    def allreduce_max(self, v):
        import torch
        t = torch.tensor([float(v)], dtype=torch.float64)
        self._dist.all_reduce(t, op=self._dist.ReduceOp.MAX)
        return float(t[0])


class ThreadComm:
    """All ranks as threads of one process (quick in-process checks at N = 2, 3, 4): make(n) returns the n endpoints."""

    class _Shared:
        def __init__(self, n):
            self.barrier = threading.Barrier(n)
            self.slots = [None] * n

    def __init__(self, shared, rank, size):
        self._s, self.rank, self.size = shared, rank, size

    @staticmethod
    def make(n):
        sh = ThreadComm._Shared(n)
        return [ThreadComm(sh, r, n) for r in range(n)]

    def _all(self, obj):
        self._s.slots[self.rank] = obj
        self._s.barrier.wait()
        out = list(self._s.slots)
        self._s.barrier.wait()
        return out

    def route(self, parcels):
        return {src: box[self.rank] for src, box in enumerate(self._all(parcels)) if self.rank in box}

    def allreduce_max(self, v):
        return max(self._all(float(v)))


# ------------------------------------------------------------------------------------------------------------------
# helpers
# ------------------------------------------------------------------------------------------------------------------
def aligned_periodic_mesh(case, ghost_planes=2):
    """Mesh whose x planes tile the periodic box exactly: spacing = L / floor(L / cutoff) (>= cutoff, so neighbour sets
    do not change), lower bound = box lower - ghost_planes * spacing on every axis. Returns (mesh, first box plane,
    box planes)."""
    from sphinxsys_b200 import hostmath as hm
    R = case.dtype
    lo, up = R(case.periodic_lower[0]), R(case.periodic_upper[0])
    L = R(up - lo)
    planes = int(np.floor(float(L) / float(case.kernel.cutoff)))
    spacing = R(L / R(planes))
    cells, lower = [], []
    for d in range(3):
        if d >= case.dim:
            cells.append(1)
            lower.append(0.0)
            continue
        lo_d = R(R(case.periodic_lower[d]) - R(ghost_planes) * spacing)
        lower.append(float(lo_d))
        # x: the box planes exactly; the other axes: cover the box plus the ghost layers (alignedPeriodicMesh, host layer)
        cells.append(planes + 2 * ghost_planes if d == 0
                     else int(np.ceil(float(R(case.periodic_upper[d]) - lo_d) / float(spacing))) + ghost_planes)
    return hm.MeshSpec(tuple(lower), float(spacing), tuple(cells)), ghost_planes, planes


def x_plane(pos, mesh):
    """Cell plane along x of every position, with the oracle's own cell arithmetic (base_mesh.hxx:9-15)."""
    if pos.shape[0] == 0:
        return np.zeros(0, dtype=np.int64)
    cell, _ = orc.cell_keys(pos, mesh)
    return cell.astype(np.int64) // (int(mesh.cells[1]) * int(mesh.cells[2]))


def plan_cuts(planes_of_particles, first, last, nranks):
    """Particle-count quantile cuts over the planes [first, last): cut[0] = first, cut[nranks] = last, every rank at
    least one plane (the host layer's planSlabCuts does the same on the GPU side)."""
    hist = np.bincount(planes_of_particles - first, minlength=last - first)[: last - first]
    cum = np.concatenate([[0], np.cumsum(hist)])
    cuts = [first]
    for r in range(1, nranks):
        target = cum[-1] * r / nranks
        c = int(np.argmin(np.abs(cum - target)))
        c = max(c, cuts[-1] - first + 1)
        c = min(c, (last - first) - (nranks - r))
        cuts.append(first + c)
    cuts.append(last)
    return cuts


# ------------------------------------------------------------------------------------------------------------------
# one rank of a decomposed run
# ------------------------------------------------------------------------------------------------------------------
class SlabRank:
    def __init__(self, case, comm, cuts, ring=False, skip_refresh=(), **oracle_kwargs):
        self.case, self.comm, self.cuts, self.ring = case, comm, list(cuts), bool(ring)
        self.skip_refresh = set(skip_refresh)  # negative tests: leave these variables stale on the ghost planes
        self.rank, self.size = comm.rank, comm.size
        self.kw = dict(oracle_kwargs)
        self.viscous = float(self.kw.get("viscosity", 0.0)) > 0.0
        self.transport = bool(self.kw.get("transport_velocity", 0))
        self.indicator = bool(self.kw.get("surface_indicator", 0))
        self.uints = INDICATOR_UINTS if self.indicator else []
        self.correction = bool(self.kw.get("correction", 0))
        self.variables = VARIABLES + (VISCOUS_VARIABLES if self.viscous else []) + (CORRECTION_VARIABLES if self.correction else [])
        self.R = np.float64 if self.kw.get("f64") else np.float32
        if len(self.cuts) != self.size + 1 or any(b <= a for a, b in zip(self.cuts, self.cuts[1:])):
            raise ValueError("cuts must be strictly increasing, one slab per rank")
        if ring:
            if not (case.periodic_axes & 1):
                raise ValueError("ring decomposition needs a case that is periodic along x")
            self.lo, self.up = self.R(case.periodic_lower[0]), self.R(case.periodic_upper[0])
            self.L = self.R(self.up - self.lo)
        elif self.cuts[0] != 0 or self.cuts[-1] != case.mesh.cells[0]:
            raise ValueError("cuts must run from 0 to the number of x planes")
        # observer probes (ObservedQuantityRecording): every rank evaluates every probe over what it stores; the rank that
        # owns the probe's cell plane has the whole neighbourhood (own planes + one ghost plane on either side) and its
        # value is the one recorded
        obs = self.kw.get("observers")
        self.observers = None if obs is None else np.asarray(obs, dtype=self.R).reshape(-1, 3)
        self.probe_series = []
        self.acoustic_steps = self.outer_steps = 0
        self.physical_time = 0.0
        self.migrated = self.wrapped = 0
        # initial state: every rank generates the case (as the GPU ranks do) and keeps its slab
        g = orc.OracleSim(case, **self.kw)
        g.exec("prepare_ck")
        pos = g.real("Position", 3).reshape(-1, 3)
        mine = self._owner(pos) == self.rank
        self.gid = np.nonzero(mine)[0].astype(np.int64)
        self.own = {nm: g.real(nm, w).reshape(-1, w)[mine].copy() for nm, w in self.variables}
        for nm in self.uints:  # registerStateVariable<int>("PreviousSurfaceIndicator", 1), "Indicator" 0
            self.own[nm] = np.full((self.gid.size, 1), 1 if nm == "PreviousSurfaceIndicator" else 0, dtype=np.uint32)
        del g
        self.sim = None
        self.rebuild(bound=False)

    # -- ownership ---------------------------------------------------------------------------------------------
    def _planes(self, pos):
        p = x_plane(np.ascontiguousarray(pos), self.case.mesh)
        if self.ring:  # a particle exactly on the upper bound (x > up is false: it is not wrapped) stays with the last rank
            p = np.clip(p, self.cuts[0], self.cuts[-1] - 1)
        return p

    def _owner(self, pos):
        return np.searchsorted(np.asarray(self.cuts), self._planes(pos), side="right") - 1

    def _names(self):
        return [nm for nm, _ in self.variables] + list(self.uints)

    def _neighbours(self):
        left, right = self.rank - 1, self.rank + 1
        if self.ring:
            return left % self.size, right % self.size
        return (left if left >= 0 else None), (right if right < self.size else None)

    # -- configuration update (SlabDecomposition::rebuild) ---------------------------------------------------------
    def rebuild(self, bound=True):
        R, case = self.R, self.case
        if self.sim is not None:  # take the own particles' state out of the stage arrays
            n = self.gid.size
            self.own = {nm: self.sim.real(nm, w).reshape(-1, w)[:n].copy() for nm, w in self.variables}
            for nm in self.uints:
                self.own[nm] = self.sim.uint(nm)[:n].copy().reshape(-1, 1)
        pos = self.own["Position"]
        if bound and case.periodic_axes:  # PeriodicBounding, axis by axis (domain_bounding.h:98-108), in Real arithmetic
            for a in range(3):
                if not (case.periodic_axes >> a & 1):
                    continue
                lo, up = R(case.periodic_lower[a]), R(case.periodic_upper[a])
                L = R(up - lo)
                x = pos[:, a]
                low, high = x < lo, x > up
                x[low] = x[low] + L
                x[high & ~low] = x[high & ~low] - L
                if a == 0:
                    self.wrapped += int(low.sum() + high.sum())
        # 1. migration: every variable of the particles whose plane belongs to another rank now
        owner = self._owner(pos)
        left, right = self._neighbours()
        away = owner != self.rank
        parcels = {}
        for dest in np.unique(owner[away]):
            if int(dest) not in (left, right):
                raise AssertionError(f"rank {self.rank}: particle migrates to rank {dest}, not a neighbour (CFL assumption)")
            sel = owner == dest
            parcels[int(dest)] = {"gid": self.gid[sel], **{nm: self.own[nm][sel] for nm in self._names()}}
        self.migrated += int(away.sum())
        arrived = self.comm.route(parcels)
        keep = ~away
        gid = [self.gid[keep]] + [box["gid"] for box in arrived.values()]
        vals = {nm: [self.own[nm][keep]] + [box[nm] for box in arrived.values()] for nm in self._names()}
        gid = np.concatenate(gid)
        order = np.argsort(gid, kind="stable")
        self.gid = gid[order]
        self.own = {nm: np.concatenate(vals[nm])[order] for nm in self._names()}
        if np.any(np.diff(self.gid) == 0):
            raise AssertionError("a particle has two owners")
        # 2. ghost planes: my first plane -> the left neighbour's right ghosts, my last plane -> the right neighbour's left ghosts
        plane = self._planes(self.own["Position"])
        if plane.size and (plane.min() < self.cuts[self.rank] or plane.max() >= self.cuts[self.rank + 1]):
            raise AssertionError("own particle outside the slab after migration")
        self.send_left = np.nonzero(plane == self.cuts[self.rank])[0]
        self.send_right = np.nonzero(plane == self.cuts[self.rank + 1] - 1)[0]
        recv = self._exchange_planes(self._names(), with_gid=True)
        self.ghost_gid = {side: recv[side]["gid"] for side in ("from_left", "from_right")}
        n_own, n_l, n_r = self.gid.size, self.ghost_gid["from_left"].size, self.ghost_gid["from_right"].size
        # 3. stage arrays: own | ghosts from the left | ghosts from the right
        local_pos = np.concatenate([self.own["Position"], recv["from_left"]["Position"], recv["from_right"]["Position"]])
        axes = case.periodic_axes & ~1 if self.ring else case.periodic_axes
        local_case = dataclasses.replace(case, fluid_pos=np.ascontiguousarray(local_pos), fluid_vel=None, periodic_axes=axes)
        sim = orc.OracleSim(local_case, **self.kw)
        for nm, w in self.variables:
            sim.real(nm, w)[:] = np.concatenate([self.own[nm], recv["from_left"][nm], recv["from_right"][nm]]).reshape(-1)
        sim.exec("set_reduce_count", n_own)
        sim.exec("cell_list_fluid")
        sim.exec("cell_list_wall")
        sim.exec("relations")
        if self.uints:
            sim.exec("surface_indication")  # sizes the indicator arrays; the values are replaced right away
            for nm in self.uints:
                sim.uint(nm)[:] = np.concatenate([self.own[nm], recv["from_left"][nm], recv["from_right"][nm]]).reshape(-1)
        self.sim, self.n_own, self.n_ghost = sim, n_own, (n_l, n_r)
        if self.observers is not None:  # fluid_observer_contact_relation.exec(); fluid_observer_pressure.writeToFile(), dambreak.cpp:223-224
            sim.exec("observer_relation")
            sim.exec("observe_pressure")
            mine = self._owner(self.observers) == self.rank
            values = np.where(mine, sim.real("Pressure", body=2).copy(), 0).astype(np.float64)
            total = np.zeros_like(values)
            for v in self.comm.route({r: values for r in range(self.size)}).values():
                total += v  # one non-zero term per probe: exact
            self.probe_series.append(total.astype(self.R))

    def _exchange_planes(self, names, with_gid=False):
        """Boundary-plane values of `names` to the neighbours; returns {"from_left": {...}, "from_right": {...}}."""
        left, right = self._neighbours()
        src = self.own if self.sim is None or with_gid else None

        def values(nm, idx):
            if src is not None:
                return src[nm][idx]
            return self.sim.real(nm, WIDTH[nm]).reshape(-1, WIDTH[nm])[idx].copy()

        parcels = {}
        for dest, idx, key, seam in ((left, self.send_left, "from_right", self.rank == 0),
                                     (right, self.send_right, "from_left", self.rank == self.size - 1)):
            if dest is None:
                continue
            box = {nm: values(nm, idx) for nm in names}
            if with_gid:
                box["gid"] = self.gid[idx]
            if self.ring and seam and "Position" in box:  # across the seam: the image position, rounded as a list entry is
                p = box["Position"].copy()
                p[:, 0] = p[:, 0] + self.L if key == "from_right" else p[:, 0] - self.L
                box["Position"] = p
            parcels.setdefault(int(dest), {})[key] = box
        got = self.comm.route(parcels)
        out = {}
        for key in ("from_left", "from_right"):
            boxes = [b[key] for b in got.values() if key in b]
            if len(boxes) > 1:
                raise AssertionError("two senders for one ghost plane")
            if boxes:
                out[key] = boxes[0]
            else:
                out[key] = {nm: np.zeros((0, WIDTH.get(nm, 1)), dtype=np.uint32 if nm in INDICATOR_UINTS else self.R) for nm in names}
                out[key]["gid"] = np.zeros(0, dtype=np.int64)
        return out

    def refresh(self, names):
        """SlabDecomposition::refreshGhosts: the named variables of the ghost planes, from their owners, in place."""
        names = [nm for nm in names if nm not in self.skip_refresh]
        recv = self._exchange_planes(names)
        n_l, n_r = self.n_ghost
        for nm in names:
            w = WIDTH[nm]
            a = self.sim.real(nm, w).reshape(-1, w)
            if recv["from_left"][nm].shape[0] != n_l or recv["from_right"][nm].shape[0] != n_r:
                raise AssertionError("ghost plane changed size between rebuilds")
            a[self.n_own:self.n_own + n_l] = recv["from_left"][nm]
            a[self.n_own + n_l:self.n_own + n_l + n_r] = recv["from_right"][nm]

    # -- one advection step (DamBreakCK::stepOuter; oracle runCK) --------------------------------------------------
    def step_outer(self):
        s, c = self.sim, self.comm
        s.exec("compression_summation")
        s.exec("density_regularization")
        s.exec("advection_setup")
        self.refresh(["VolumetricMeasure"])
        # viscous force, kernel gradient integral, transport correction (order of lid_driven_cavity_sycl.cpp:268-276): they
        # read Position, VolumetricMeasure and Velocity of the neighbours, all current on the ghost planes at this point
        # (velocities have not changed since the configuration update), so no further refresh is needed
        if self.viscous:
            s.exec("viscous_force")
        if self.transport:
            s.exec("kernel_gradient_integral")
            s.exec("transport_velocity_correction", 1, 0)
        adv_dt = s.exec("advection_dt_of", c.allreduce_max(s.exec("advection_dt_reduced")))
        if self.indicator:  # fluid_boundary_indicator.exec(), dambreak.cpp:192: two sweeps, the second reads the first's result
            s.exec("surface_interact")
            # the GPU runs the sweep on the own particles only; ghosts close to the cut would get the right value here by
            # accident (their neighbourhood is complete), which would hide a missing refresh: wipe what was computed there
            s.real("PositionDivergence")[self.n_own:] = 0
            self.refresh(["PositionDivergence"])
            s.exec("surface_update")
        if self.correction:  # LinearCorrectionMatrix<Inner<WithUpdate>, Contact<>>: both half steps read B of the neighbours
            s.exec("linear_correction")
            self.refresh(["LinearCorrectionMatrix"])
        relax = 0.0
        while relax < adv_dt:
            dt = s.exec("acoustic_dt_of", c.allreduce_max(s.exec("acoustic_dt_reduced")))
            s.exec("acoustic1_init", dt)
            self.refresh(["Pressure"])
            s.exec("acoustic1_inner")
            s.exec("acoustic1_wall")
            s.exec("acoustic1_update", dt)
            self.refresh(["Velocity"])
            s.exec("acoustic2", dt)
            relax += dt
            self.physical_time += dt
            self.acoustic_steps += 1
        s.exec("update_position")
        self.outer_steps += 1
        self.rebuild()

    def recut(self, plan, limit):
        """SlabDecomposition::recut: new cuts from the CURRENT particles-per-plane histogram of all ranks
        (`plan(hist, nranks)`, the host layer's planSlabCuts), every cut moved at most to the far end of an adjacent slab
        (`limit(old, wanted)`, limitCutMoves), so that the hand-over stays between neighbour ranks — which rebuild()
        asserts. Returns True if the cuts changed."""
        # a ring is cut over the planes of the periodic box only; its seam (first and last cut) stays where it is
        first = self.cuts[0] if self.ring else 0
        planes_total = (self.cuts[-1] - first) if self.ring else int(self.case.mesh.cells[0])
        mine = np.bincount(self._planes(self.sim.real("Position", 3).reshape(-1, 3)[: self.n_own]) - first, minlength=planes_total)
        hist = np.zeros(planes_total, dtype=np.uint64)
        for h in self.comm.route({r: mine for r in range(self.size)}).values():
            hist += h.astype(np.uint64)
        wanted = np.asarray(plan(hist, self.size), dtype=np.int32) + np.int32(first)
        new = [int(c) for c in limit(np.asarray(self.cuts, dtype=np.int32), wanted)]
        changed = new != self.cuts
        self.cuts = new
        self.rebuild(bound=False)
        return changed

    def own_state(self):
        """{name: values of the own particles} plus "gid" (ascending), as of the last configuration update."""
        out = {nm: self.sim.real(nm, w).reshape(-1, w)[: self.n_own].copy() for nm, w in self.variables}
        for nm in self.uints:
            out[nm] = self.sim.uint(nm)[: self.n_own].copy().reshape(-1, 1)
        out["gid"] = self.gid.copy()
        return out


def run_threads(case, nranks, cuts, steps, ring=False, skip_refresh=(), **oracle_kwargs):
    """All ranks in this process (threads); returns the per-rank own_state() list and the rank objects."""
    comms = ThreadComm.make(nranks)
    ranks, errors = [None] * nranks, []

    def work(r):
        try:
            sr = SlabRank(case, comms[r], cuts, ring=ring, skip_refresh=skip_refresh, **oracle_kwargs)
            for _ in range(steps):
                sr.step_outer()
            ranks[r] = sr
        except BaseException as e:  # noqa: BLE001 - reported by the caller; the barrier is broken so the others stop too
            errors.append((r, e))
            comms[r]._s.barrier.abort()

    threads = [threading.Thread(target=work, args=(r,)) for r in range(nranks)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    real = [e for e in errors if not isinstance(e[1], threading.BrokenBarrierError)]
    if real or errors:
        raise (real or errors)[0][1]
    return [sr.own_state() for sr in ranks], ranks


def gather_by_gid(states, n_total):
    """Per-rank own states -> global arrays in the single-domain particle numbering."""
    seen = np.zeros(n_total, dtype=np.int64)
    out = {}
    for st in states:
        seen[st["gid"]] += 1
    if not np.all(seen == 1):
        raise AssertionError(f"ownership is not a partition: {int((seen == 0).sum())} lost, {int((seen > 1).sum())} duplicated")
    for nm in [k for k in states[0] if k != "gid"]:
        w = WIDTH.get(nm, 1)
        a = np.empty((n_total, w), dtype=states[0][nm].dtype)
        for st in states:
            a[st["gid"]] = st[nm]
        out[nm] = a
    return out
