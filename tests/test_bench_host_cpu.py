"""CPU tests of the measurement harness (bench.py) and of the host-side bookkeeping that travels with the N-GPU numbers."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_particle_digest_is_order_and_partition_independent_and_bit_sensitive():
    """bench.py's `parity` record compares ONE 64-bit word per run: the digest must not depend on storage order or on how the
    particles are split over ranks, and must change when a single bit of a single particle changes."""
    from sphinxsys_b200.host import particle_digest
    rng = np.random.default_rng(3)
    n = 50_000
    ids = np.arange(n, dtype=np.uint32)
    pos = rng.standard_normal((n, 4)).astype(np.float32)  # device layout: the 4th float is padding and must not count
    vel = rng.standard_normal((n, 4)).astype(np.float32)
    with np.errstate(over="ignore"):
        whole = particle_digest(ids, pos, vel)
        p = rng.permutation(n)
        cut = [0, 7, 12_345, 30_000, n]
        parts = sum(particle_digest(ids[p][a:b], pos[p][a:b], vel[p][a:b]) for a, b in zip(cut[:-1], cut[1:])) % (1 << 64)
        assert parts == whole
        pad = pos.copy()
        pad[:, 3] = 123.0
        assert particle_digest(ids, pad, vel) == whole
        one_ulp = vel.copy()
        one_ulp[n // 2, 2] = np.nextafter(one_ulp[n // 2, 2], np.float32(10.0))
        assert particle_digest(ids, pos, one_ulp) != whole
        swapped = pos.copy()
        swapped[[1, 2]] = swapped[[2, 1]]  # same multiset of positions, different owners
        assert particle_digest(ids, swapped, vel) != whole
        assert particle_digest(ids[:0], pos[:0], vel[:0]) == 0


def test_reference_arm_uses_all_cores_under_torchrun_environment():
    """torchrun exports OMP_NUM_THREADS=1; the reference arm must still use every core of the affinity mask and must keep the
    case it is asked for (same spacing at every N), bounding only the number of steps."""
    env = dict(os.environ, OMP_NUM_THREADS="1", RANK="0", WORLD_SIZE="2", LOCAL_RANK="0")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "2",
                        "--warmup", "1", "--ref-dp", "0.05"], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    cores = len(os.sched_getaffinity(0))
    assert line["impl"] == "reference" and line["n_gpus"] == 2
    assert line["cpu_baseline"]["cores"] == cores, line["cpu_baseline"]
    assert line["config"]["n_fluid_global"] == 8000 and "dp=0.05" in line["config"]["workload"]
    assert line["steps"] == 2 and line["value"] > 0 and line["e2e"]["h2d_bytes_per_step"] == 0
    # the other ranks print nothing and exit 0
    env["RANK"] = "1"
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "2",
                        "--warmup", "1", "--ref-dp", "0.05"], capture_output=True, text=True, env=env, timeout=60)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_second_roofline_arithmetic():
    """issue / L1 fractions: counters per particle x particles over the slots of the measured launch time."""
    sys.path.insert(0, ROOT)
    import bench
    counters = bench.kernel_counters()
    assert {"k_a2", "k_a1_interact"} <= set(counters)
    n, ms, mhz = 4_096_000, 0.75, 1965.0
    rec = bench.second_roofline("k_a2", n, ms, mhz, counters)
    cycles = ms * 1e-3 * mhz * 1e6
    assert abs(rec["issue_frac"] - counters["k_a2"]["warp_instructions_per_particle"] * n / (cycles * 148 * 4)) < 1e-12
    assert abs(rec["l1_data_pipe_frac"] - counters["k_a2"]["l1_wavefronts_per_particle"] * n / (cycles * 148)) < 1e-12
    assert 0.3 < rec["issue_frac"] < 1.0 and 0.3 < rec["l1_data_pipe_frac"] < 1.0
    assert bench.second_roofline("no_such_kernel", n, ms, mhz, counters) is None
