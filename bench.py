#!/usr/bin/env python
"""bench.py -- `fix eph` hot path throughput on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # our arm (CUDA path through the C ABI)
    python bench.py --impl reference --gpus N --steps K ...  # reference arm: the reference's own CPU fix

Workload: fcc Ni, n^3 unit cells (n = 100 -> 4 000 000 atoms), a = 3.52 A, N(0, 0.05 A)
displacements, Maxwell velocities at 600 K plus one 10 keV primary knock-on atom, flags 7 (friction + random + FDM),
model 4, `Data/Ni/Ni_PRB2019.beta` (shipped copy), FDM grid 64^3, full neighbour list at r_c + 2 A = 7 A, time step of
the cascade's adaptive rule while the PKA is fast (1e-3 A / v_max = 5.5e-7 ps).

One step is what LAMMPS' Verlet loop makes the fix do: initial_integrate -> (LAMMPS: ghost refresh; every 10th step
re-neighbouring: the fix registers its atoms again, receives the full list LAMMPS built for it -- device-resident, as
from LAMMPS-KOKKOS; `--neigh device` makes the engine build it from the positions instead -- and rebuilds its inner
list) -> post_force -> final_integrate -> end_of_step.  The atoms move and all of it is inside the timed region
(`--mode static` times post_force + end_of_step on frozen atoms instead, the round-1 measurement).  `value` has everything resident in HBM; `e2e` goes through the same C ABI with pinned HOST
buffers.  On N > 1 GPUs the engine's own data plane runs the ghost exchanges (NVLink peer memory, or NCCL send/receive),
the source all-reduce (NCCL) and the grid solve, and after the timed steps the result is checked atom by atom against
the whole box on one GPU (`parity_vs_n1`); `step_breakdown` says where the step goes.
Prints ONE JSON line.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "user-eph_b200"))
sys.path.insert(0, ROOT)

METRIC = "fix eph atom-steps/s (Ni 4M atoms)"
UNIT = "atom-steps/s"
REBUILD_EVERY = 10
V_PKA = 1813.0                      # A/ps: 10 keV Ni (Tests/MD_Run/run.lmp:69-73)
DT = 1e-3 / V_PKA                   # ps: the adaptive rule's step while the PKA is fast (Tests/MD_Run/run.lmp:94-97, :177-189)
MASS = 58.71
FTM2V = 1.0 / 1.0364269e-4
CUTOFF = 7.0                        # r_c + skin


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--mode", default="trajectory", choices=["trajectory", "static"],
                    help="trajectory (default): integrator hooks + re-neighbouring every 10 steps inside the timed region; "
                         "static: post_force + end_of_step on frozen atoms (steady state of the two sweeps)")
    ap.add_argument("--cells", type=int, default=100, help="fcc unit cells per box edge (100 -> 4M atoms)")
    ap.add_argument("--grid", type=int, default=64, help="FDM grid points per edge")
    ap.add_argument("--dt", type=float, default=DT)
    ap.add_argument("--neigh", default="lammps", choices=["lammps", "device"],
                    help="where the full neighbour list comes from at a re-neighbouring: lammps (default) = handed over as a "
                         "device-resident CSR, the way GPU-resident LAMMPS (KOKKOS) provides the list its fix requested -- like the "
                         "reference fix, which is given its list by LAMMPS; device = built by the engine from the positions "
                         "(fix keyword `neigh device`: LAMMPS then builds no list for this fix)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-fdm-bench", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the C4 (4-element) leg and the small-box check against the reference")
    ap.add_argument("--no-check", action="store_true", help="multi-GPU: skip the atom-by-atom comparison with the whole box on one GPU")
    ap.add_argument("--overlap", action="store_true", help="multi-GPU: ghost exchange on a second stream behind a boundary-first density pass")
    ap.add_argument("--sharded-grid", action="store_true", help="multi-GPU: every rank advances only its z-slab of the FDM grid "
                    "(halo planes between sub-steps, all-gather at the end) instead of solving the whole grid redundantly; "
                    "needs --grid divisible by --gpus")
    ap.add_argument("--weak", action="store_true", help="weak scaling (supplementary): --cells^3 unit cells and --grid^3 grid "
                    "cells PER GPU (config C5 at 8 GPUs: 32 M atoms); the default is strong scaling of the named 4 M-atom box")
    ap.add_argument("--elements", type=int, default=1, help="config C4: 4 = the NiCoCrFe table with types uniform random")
    ap.add_argument("--cpu-cells", type=int, default=32, help="edge of each CPU replica (32 -> 131 072 atoms)")
    ap.add_argument("--cpu-steps", type=int, default=8, help="timed steps of every CPU-baseline replica (about 12 s of work per core)")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------------------------
# workload
# ---------------------------------------------------------------------------------------------------------------
def beta_file(elements):
    from eph_harness import harness as H
    return H.shipped_beta("Ni_PRB2019" if elements == 1 else "NiCoCrFe_PRB2019")


def build_workload(cells, brick=None, elements=1, with_list=False):
    from eph_harness import harness as H
    s = H.make_system(cells, brick=brick, ntypes=elements, with_list=with_list)
    # the primary knock-on atom of config C3: 10 keV along (0.835, 0.544, 0.082) (Tests/MD_Run/run.lmp:69-73); tag 1
    pka = np.nonzero(s["tag"] == 1)[0]
    s["v"][pka] = V_PKA * np.array([0.835115, 0.543981, 0.081652])
    return s


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index = index
        self.rows = []
        self.stop_flag = False
        self.proc = None

    def run(self):
        try:
            p = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                  "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            return
        self.proc = p
        for line in p.stdout:
            self.rows.append(line.strip())
            if self.stop_flag:
                break
        p.kill()

    def summary(self):
        self.stop_flag = True
        time.sleep(0.15)
        try:
            self.proc.kill()
        except Exception:
            pass
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            c = [t.strip() for t in r.split(",")]
            try:
                sm.append(float(c[0]))
                mx.append(float(c[1]))
            except Exception:
                continue
            for n, v in zip(names, c[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------------------
# CPU baseline: the reference's own fix (compiled, unmodified) or the oracle port, as independent replicas
# ---------------------------------------------------------------------------------------------------------------
_cpu_state = {}


def _cpu_init(cells, grid, dt, seed_base):
    """one replica per worker process, built once: the unmodified reference fix on its own periodic box"""
    import multiprocessing as mp
    sys.path.insert(0, os.path.join(ROOT, "user-eph_b200"))
    sys.path.insert(0, ROOT)
    from eph_harness import harness as H
    k = mp.current_process()._identity[0] if mp.current_process()._identity else 0
    s = H.make_system(cells, pos_seed=1234 + k, vel_seed=101 + k)
    s["v"][0] = V_PKA * np.array([0.835115, 0.543981, 0.081652])
    s["v"][s["nlocal"]:][s["ghost_owner"] == 0] = s["v"][0]
    xi = np.random.default_rng(seed_base + k).normal(size=(s["nlocal"], 3))
    box = [0, s["box"][0], 0, s["box"][1], 0, s["box"][2]]
    from oracle import reference as R
    if R.available():
        devnull = os.open(os.devnull, os.O_WRONLY)   # the reference prints a banner per fix
        saved = os.dup(1)
        os.dup2(devnull, 1)
        try:
            drv = R.fix_driver(s, H.fix_args(7, beta_file(1), ["Ni"], grid=(grid,) * 3), dt=dt)
        finally:
            os.dup2(saved, 1)

        def step():
            drv.set_xi(xi)
            drv.post_force()
            drv.end_of_step()
        kind = "reference"
    else:
        from oracle import oracle as O
        fx = O.Fix(s, O.Beta(path=beta_file(1)), O.FDM(grid, grid, grid, box, 300.0, 3.5e-6, 1.0, 0.1248), 7, dt=dt)

        def step():
            fx.post_force(xi)
            fx.end_of_step()
        kind = "port"
    step()                                           # untimed first step
    _cpu_state.update(step=step, natoms=s["nlocal"], kind=kind)


def _cpu_run(nsteps):
    t0 = time.perf_counter()
    for _ in range(nsteps):
        _cpu_state["step"]()
    return _cpu_state["natoms"] * nsteps, time.perf_counter() - t0, _cpu_state["kind"]


class CpuReplicas:
    """`procs` concurrent single-rank replicas of the reference fix (it has no threads and MPI is not installed, so
    replicas stand in for ranks), each on its own 4 * cells^3-atom periodic box with the
    workload's grid, parametrisation, PKA and time step."""

    def __init__(self, cells, grid, dt, procs=None):
        import multiprocessing as mp
        self.procs = procs or max(1, (os.cpu_count() or 1))
        self.cells, self.grid = cells, grid
        self.pool = mp.get_context("spawn").Pool(self.procs, initializer=_cpu_init, initargs=(cells, grid, dt, 777))
        self.pool.map(_cpu_run, [0] * self.procs)   # every worker has built its replica

    def sample(self, nsteps):
        t0 = time.perf_counter()
        res = self.pool.map(_cpu_run, [nsteps] * self.procs, chunksize=1)
        wall = time.perf_counter() - t0
        work, tmax = sum(r[0] for r in res), max(r[1] for r in res)
        return {"value": work / tmax, "unit": UNIT, "cores": self.procs, "kind": res[0][2],
                "sample": "%d concurrent single-rank replicas of %d Ni atoms (n=%d) with a %d^3 grid, %d timed steps each, "
                          "flags 7, model 4, Ni_PRB2019; slowest replica %.2f s, wall %.1f s"
                          % (self.procs, 4 * self.cells ** 3, self.cells, self.grid, nsteps, tmax, wall)}

    def close(self):
        self.pool.close()
        self.pool.join()


def run_reference_arm(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    rep = CpuReplicas(a.cpu_cells, a.grid, a.dt)
    per_step, base = [], None
    for k in range(a.warmup + a.steps):
        base = rep.sample(1)
        if k >= a.warmup:
            per_step.append(base["value"])
    rep.close()
    value = float(np.mean(per_step))
    natoms = 4 * a.cells ** 3
    base["value"] = value
    out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
           "ms_per_step": 1e3 * natoms / value, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
           "dtype": "f64", "data": "synthetic", "impl": "reference",
           "config": workload_config(a, natoms), "cpu_baseline": base,
           "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "note": "the unmodified reference fix (post_force + end_of_step) on all host cores; every step is a bounded sample "
                   "(one replica of %d atoms with the workload's %d^3 grid per core, one fix step each); ms_per_step is the "
                   "time the %d-atom workload would need at that rate" % (4 * a.cpu_cells ** 3, a.grid, natoms)}
    print(json.dumps(out))


def workload_config(a, natoms):
    multi = getattr(a, "elements", 1) > 1
    beta = "Data/NiCoCrFe/NiCoCrFe_PRB2019.beta" if multi else "Data/Ni/Ni_PRB2019.beta"
    steps = ("trajectory: initial_integrate, post_force, final_integrate, end_of_step on moving atoms, re-neighbouring every %d "
             "steps (atoms registered again, %s, inner list rebuilt), all inside the timed region"
             % (REBUILD_EVERY, "full list handed over as a device-resident CSR like LAMMPS-KOKKOS does" if a.neigh == "lammps"
                else "full list rebuilt by the engine from the positions")) if a.mode == "trajectory" else "static: post_force + end_of_step on frozen atoms"
    if getattr(a, "weak", False) and a.gpus > 1:
        return {"workload": "C5-style weak scaling: Ni fcc, %d^3 cells (= %d atoms) and a %d^3 grid per GPU, %d atoms in all, "
                            "flags 7, model 4, dt %.3g ps, full list at 7 A" % (a.cells, 4 * a.cells ** 3, a.grid, natoms, a.dt),
                "atoms": natoms, "fdm_grid_per_gpu": [a.grid] * 3, "beta_file": beta, "step": steps,
                "l2": "inputs far larger than the 126 MB L2; no flush needed", "parallelism": "spatial bricks, one rank per GPU"}
    name = "C4: NiCoCrFe fcc alloy (types uniform random)" if multi else "C3: Ni fcc"
    return {"workload": "%s %d^3 cells = %d atoms, one 10 keV PKA, flags 7 (friction+random+FDM), model 4, "
                        "FDM grid %d^3, dt %.3g ps, full list at 7 A" % (name, a.cells, natoms, a.grid, a.dt),
            "atoms": natoms, "fdm_grid": [a.grid] * 3, "beta_file": beta + " (shipped copy, tests/golden/data)", "step": steps,
            "l2": "inputs (neighbour list + per-atom arrays) far larger than the 126 MB L2; no flush needed",
            "parallelism": "spatial bricks, one rank per GPU; ghost exchanges (peer memory or NCCL send/receive), source all-reduce (NCCL) and grid solve by the engine"}


# ---------------------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------------------
class Verlet:
    """Plays LAMMPS' Verlet loop around the fix for one rank, everything on the device: the fix hooks are the engine's
    C-ABI calls, LAMMPS' own part (ghost positions and velocities follow their owners, force_clear, re-neighbouring
    cadence) is done here; the ghost refresh itself is the engine's eph_b200_refresh_ghosts (what Comm::forward_comm() with
    ghost_velocity does, for x and v that live on the device; over NCCL between ranks)."""

    def __init__(self, eng, s, plan, dist, dev, a, torch):
        self.eng, self.s, self.plan, self.dist, self.torch, self.a = eng, s, plan, dist, torch, a
        nl, ng = s["nlocal"], s["nghost"]
        t = lambda arr, dt: torch.as_tensor(np.ascontiguousarray(arr), dtype=dt, device=dev)
        self.nl, self.ng = nl, ng
        self.x, self.v = t(s["x"], torch.float64), t(s["v"], torch.float64)
        self.f = torch.zeros((nl, 3), dtype=torch.float64, device=dev)
        self.type, self.mask, self.tag = t(s["type"], torch.int32), t(s["mask"], torch.int32), t(s["tag"], torch.int64)
        self.owner = t(plan.self_owner, torch.int32)
        self.mass = np.array([0.0] + [MASS] * s["ntypes"])
        self.dtv, self.dtf = a.dt, 0.5 * a.dt * FTM2V
        self.world = dist.get_world_size() if dist else 1
        self.csr = None
        self.register(first=True)
        if a.neigh == "lammps":
            # LAMMPS' list, device-resident: built once here with the engine's own kernel (the atoms of this workload move
            # by less than 1e-4 A between re-neighbourings, so every rebuild would give the same rows)
            off, ne = eng.get_neighbors()
            self.csr = (torch.as_tensor(off, device=dev), torch.as_tensor(ne, device=dev))

    def refresh_ghosts(self):
        """LAMMPS' forward comm of x and v (comm->ghost_velocity is set by the fix, fix_eph.cpp:82), on the device: images of
        the rank's own atoms are shifted copies, ghosts owned by other ranks arrive over the engine's NCCL ghost map"""
        self.eng.refresh_ghosts(self.x, self.v)

    def register(self, first=False):
        """what FixEPHB200::upload_topology does when LAMMPS has re-neighboured: atoms, full list (built on the device from
        the positions: `neigh device`), ghost map"""
        eng = self.eng
        eng.set_atoms(self.nl, self.ng, self.type, self.mask, self.tag, self.owner)
        if self.csr is not None:
            eng.set_neighbors(*self.csr)
        else:
            eng.build_neighbors(self.x, CUTOFF)
        if self.world > 1:
            eng.set_ghost_map(self.plan)
        eng.refresh_ghosts(self.x, self.v)     # first call after set_atoms: records the image shifts of the fresh ghosts

    def step(self, k):
        eng = self.eng
        if self.a.mode == "static":
            eng.post_force(self.x, self.v, self.f, None, k)
            eng.end_of_step(None, self.v, want_energy=False)
            return
        eng.initial_integrate(self.x, self.v, self.f, self.mass, self.dtv, self.dtf)
        self.refresh_ghosts()
        if k % REBUILD_EVERY == 0:
            self.register()
        self.f.zero_()                      # LAMMPS: force_clear; there is no pair style in this workload
        eng.post_force(self.x, self.v, self.f, None, k)
        eng.final_integrate(self.v, self.f, self.mass, self.dtf)
        eng.end_of_step(None, self.v, want_energy=False)


def make_engine(lib, host, a, s, gridn, box, local, rank, world, stream, elements=None):
    elements = elements or a.elements
    eng = lib.Engine(list(range(elements)), flags=7, seed=12345, device=local, rank=rank, nranks=world, stream=stream)
    eng.set_tables_from(host.BetaTables(path=beta_file(elements)))
    eng.set_grid(gridn[0], gridn[1], gridn[2], box, 300.0, 1.0, 3.5e-6, 0.1248)
    eng.set_dt(a.dt)
    eng.set_skin(2.0)
    return eng


def run_b200(a):
    import torch
    import torch.distributed as dist
    from eph_b200 import host, lib
    from eph_harness import parallel as P

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != a.gpus:
        raise SystemExit("bench.py: --gpus %d but WORLD_SIZE %d (launch N > 1 with torch.distributed.run)" % (a.gpus, world))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    D = dist if world > 1 else None

    grid = P.brick_grid(world)
    cells = tuple(a.cells * g for g in grid) if (a.weak and world > 1) else a.cells
    gridn = tuple(a.grid * g for g in grid) if (a.weak and world > 1) else (a.grid,) * 3
    s = build_workload(cells, brick=(rank, grid) if world > 1 else None, elements=a.elements)
    if world == 1:
        s["grid"] = (1, 1, 1)
    nl, ng = s["nlocal"], s["nghost"]
    natoms = s["natoms"]
    box = [0, s["box"][0], 0, s["box"][1], 0, s["box"][2]]
    plan = P.ExchangePlan(s, rank, world, D)

    # a dedicated stream shared by torch and the engine, so torch's CUDA events bracket the engine's kernels
    tstream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(tstream)
    stream = tstream.cuda_stream
    assert stream != 0
    eng = make_engine(lib, host, a, s, gridn, box, local, rank, world, stream)
    sharded = bool(D) and a.sharded_grid and P.grid_slab(gridn[2], rank, world) is not None
    gstream = cstream = None
    if D:
        # the engine's own data plane: NCCL communicator (its 128-byte id travels over torch.distributed), then plain
        # post_force / end_of_step do the ghost exchange, the source all-reduce and the grid solve
        P.attach_comm(eng, D, rank, world)
        eng.set_grid_sharding(sharded)
        gstream = torch.cuda.Stream(device=dev)      # source all-reduce + solve overlap the next step's density pass
        eng.set_grid_stream(gstream.cuda_stream)
        if a.overlap:
            cstream = torch.cuda.Stream(device=dev)
            eng.set_comm_stream(cstream.cuda_stream)
    md = Verlet(eng, s, plan, D, dev, a, torch)
    n_nb = float(eng.get_neighbors_count()) / max(nl, 1)

    def timed(fn, steps, first):
        """barrier + synchronize on both sides, device time via CUDA events, max over ranks"""
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if D:
            D.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        e0.record()
        for k in range(steps):
            fn(first + k)
        e1.record()
        if D:
            D.barrier()
        torch.cuda.synchronize()
        ms = max(e0.elapsed_time(e1), 0.0)
        wall = 1e3 * (time.perf_counter() - t0)
        if D:
            tt = torch.tensor([ms, wall], dtype=torch.float64, device=dev)
            D.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms, wall = float(tt[0]), float(tt[1])
        return ms, wall

    # warm-up steps 1 .. W, timed steps W+1 .. W+K: the re-neighbouring steps (multiples of 10) fall where they fall
    for k in range(1, a.warmup + 1):
        md.step(k)
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    # CUDA events around the two list sweeps only inside the timed region (the roofline's kernel durations): a pair of
    # events costs about 3 us of stream time, which for all 13 kernels of a step is 9 % of a 500 k-atom brick's step
    eng.set_profiling(2)
    l0 = eng.launch_count()
    stats0 = eng.list_stats()
    ms, _ = timed(md.step, a.steps, a.warmup + 1)
    launches = eng.launch_count() - l0
    ktimes = eng.kernel_times()
    stats1 = eng.list_stats()
    clocks = sampler.summary() if rank == 0 else None
    if eng.status_word() & 0x100:
        raise SystemExit("bench.py: a peer-memory ghost exchange timed out on rank %d (results invalid)" % rank)
    # the other kernels: the same number of further steps with every launch bracketed, outside the timed region
    eng.set_profiling(1)
    ms_all, _ = timed(md.step, a.steps, a.warmup + a.steps + 1)
    ktimes_all = eng.kernel_times()
    eng.set_profiling(0)
    ktimes = dict(ktimes_all, **ktimes)
    ms_per_step = ms / a.steps
    value = natoms * a.steps / (ms * 1e-3)
    rebuilds = len([k for k in range(a.warmup + 1, a.warmup + a.steps + 1) if k % REBUILD_EVERY == 0]) if a.mode == "trajectory" else 0

    # ---- roofline of the dominant kernel (SURVEY.md 8d per-sweep algorithmic bytes) ----
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    # rho sweep 36+4N, w sweep 84+4N, f sweep 180+4N bytes per atom.  density_sweep does the work of the
    # reference's rho AND w sweeps, force_sweep that of its f sweep (friction + random).
    alg = {"density_sweep": (36 + 4 * n_nb) + (84 + 4 * n_nb), "force_sweep": 180 + 4 * n_nb}
    per_kernel = {k: {"ms_avg": v[0] / max(v[1], 1), "launches": v[1], "ms_per_step": v[0] / a.steps} for k, v in ktimes.items()}
    dom = max((k for k in per_kernel if k in alg), key=lambda k: per_kernel[k]["ms_avg"], default=None)
    roofline = None
    if dom:
        ach = alg[dom] * nl / (per_kernel[dom]["ms_avg"] * 1e-3) / 1e9
        b_step = 420 + 12 * n_nb
        # DRAM bytes per launch of the same kernel on the same workload from the committed ncu capture (tools/ncu_traffic.py)
        traffic, traffic_src = None, None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "r2_traffic_4M.json")))
            if tj["atoms"] == nl and dom in tj["kernels"]:
                traffic, traffic_src = tj["kernels"][dom]["dram_bytes"], "profiles/r2_traffic_4M.json (%s)" % tj["report"]
        except Exception:
            pass
        roofline = {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                    "traffic": traffic, "traffic_unit": "DRAM bytes per launch (ncu)", "traffic_source": traffic_src,
                    "algorithmic_bytes_per_launch": alg[dom] * nl, "peak_source": peak_src,
                    "algorithmic_bytes_per_atom": alg[dom], "kernel_ms": per_kernel[dom]["ms_avg"],
                    "step": {"algorithmic_bytes_per_atom_step": b_step, "achieved": b_step * value / 1e9 / world,
                             "frac": b_step * value / 1e9 / world / peak, "note": "whole step as timed (incl. integrator hooks and list maintenance), per GPU"},
                    "kernels_ms": {k: round(v["ms_avg"], 4) for k, v in per_kernel.items()},
                    "kernels_ms_per_step": {k: round(v["ms_per_step"], 4) for k, v in per_kernel.items()},
                    "kernels_note": "the two sweeps: CUDA events on the launch stream inside the timed region; the other kernels: the %d steps "
                                    "that follow it, every launch bracketed (%.4f ms per step with all those events)" % (a.steps, ms_all / a.steps)}

    # ---- where the step goes (rank 0's kernels; what the kernels on the main stream do not account for is launch gaps,
    # registration and waiting for neighbours) ----
    breakdown = None
    if roofline:
        kps = roofline["kernels_ms_per_step"]
        side = ("source_allreduce", "source_reduce_scatter", "grid_allgather", "grid_halo", "fdm_substep") if D else ()
        groups = {"sweeps": ("density_sweep", "force_sweep", "density_sweep_boundary"),
                  "ghost_exchanges": ("pack_payload", "ghost_exchange", "unpack_payload", "refresh_send", "refresh_recv", "ghost_images"),
                  "list_maintenance": ("inner_list_build", "build_neighbors"),
                  "stand_by_launches": ("density_sweep_fallback", "force_sweep_fallback")}
        named = set(sum(groups.values(), ())) | set(side)
        parts = {g: round(sum(kps.get(k, 0.0) for k in ks), 4) for g, ks in groups.items()}
        parts["streaming_kernels"] = round(sum(v for k, v in kps.items() if k not in named), 4)
        parts["gaps_registration_and_waiting"] = round(ms_per_step - sum(parts.values()), 4)
        non_sweep = {k: v for k, v in parts.items() if k != "sweeps"}
        breakdown = {"ms_per_step": parts, "largest_non_sweep_part": max(non_sweep, key=non_sweep.get),
                     "on_the_grid_stream_ms_per_step": {k: kps[k] for k in side if k in kps},
                     "note": "rank 0, main stream; the all-reduce of the source term and the grid solve run on a second stream under the next step's density pass"}

    # ---- multi-GPU: the same steps on the whole box on one GPU, atom by atom ----
    parity = None
    if D and not a.no_check:
        parity = check_against_one_gpu(a, torch, dist, lib, host, P, eng, md, s, gridn, box, cells, local, rank, world, dev, stream)

    # ---- end to end through the C ABI with pinned host buffers ----
    e2e = None
    if not a.no_e2e:
        e2e = run_e2e(a, torch, eng, md, s, plan, D, timed, natoms, world)
        md.register()

    # ---- extras on one GPU: FDM micro-benchmarks, the 4-element configuration, a small box against the reference ----
    fdm = extras = None
    if rank == 0 and world == 1:
        if not a.no_fdm_bench:
            fdm = fdm_bench(lib, host, stream, local, peak)
        if not a.no_extras:
            extras = {"C4_NiCoCrFe": c4_leg(a, torch, lib, host, P, s, gridn, box, local, stream, dev, natoms),
                      "parity_vs_reference": small_box_check(a, torch, lib, host, local, stream, dev)}

    if rank == 0:
        out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
               "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak" if a.weak else "strong", "vs_baseline": None, "dtype": "f64",
               "data": "synthetic",
               "config": dict(workload_config(a, natoms), mean_list_length=n_nb, ghosts_rank0=ng, brick_grid=list(grid),
                              grid_solve="z-slabs + halo planes + all-gather" if sharded else "whole grid on every rank",
                              exchange_bytes_per_step_rank0=(eng.exchange_bytes if D else 0), records=eng.precision(),
                              ghost_exchange_transport={0: None, 1: "grouped ncclSend/ncclRecv", 2: "NVLink peer memory (one kernel stores the rows into the receivers' windows)"}[eng.comm_transport()],
                              reneighbourings_in_timed_region=rebuilds,
                              list_stats={k: stats1[k] - stats0[k] for k in stats1}),
               "clocks": clocks, "gpu_launches": int(launches), "roofline": roofline, "step_breakdown": breakdown, "e2e": e2e, "fdm": fdm}
        if parity is not None:
            out["parity_vs_n1"] = parity
        if extras is not None:
            out["extras"] = extras
        if not a.no_cpu_baseline and world == 1:
            rep = CpuReplicas(a.cpu_cells, a.grid, a.dt)
            out["cpu_baseline"] = rep.sample(a.cpu_steps)
            rep.close()
        print(json.dumps(out), flush=True)
    if D:
        D.barrier()
        dist.destroy_process_group()


def check_against_one_gpu(a, torch, dist, lib, host, P, eng, md, s, gridn, box, cells, local, rank, world, dev, stream, nsteps=3):
    """`nsteps` further trajectory steps on the bricks, then the same steps from the same state on the WHOLE box on rank
    0's GPU alone; forces, densities, positions and the grid are compared atom by atom / cell by cell."""
    natoms = s["natoms"]
    nl = s["nlocal"]
    tag0 = md.tag[:nl].long() - 1
    # state of all atoms before the checked steps, gathered by tag on rank 0
    def gather(t3):
        g = torch.zeros((natoms, t3.shape[1]), dtype=torch.float64, device=dev)
        g[tag0] = t3[:nl]
        dist.reduce(g, dst=0)
        return g
    x0, v0, f0 = gather(md.x), gather(md.v), gather(md.f)
    T0 = torch.as_tensor(eng.get_grid(0))
    k0 = 100000
    for k in range(nsteps):
        md.step(k0 + 1 + k)       # no re-neighbouring inside (k0 + 1 .. k0 + 3)
    rho_b = torch.as_tensor(eng.probe(0)[:nl], device=dev).reshape(-1, 1)
    xb, fb, rb = gather(md.x), gather(md.f), gather(rho_b)
    Tb = eng.get_grid(0)
    res = None
    if rank == 0:
        from eph_harness import harness as H
        w = H.make_system(cells, ntypes=a.elements, with_list=False)
        order = torch.as_tensor(w["tag"][: w["nlocal"]] - 1, device=dev)
        w["x"][: w["nlocal"]] = x0[order].cpu().numpy()
        w["v"][: w["nlocal"]] = v0[order].cpu().numpy()
        own = w["ghost_owner"]
        shift = H.make_system.__globals__["np"].zeros(0)
        plan1 = P.ExchangePlan(w, 0, 1)
        a1 = argparse.Namespace(**vars(a))
        e1 = make_engine(lib, host, a1, w, gridn, box, local, 0, 1, stream)
        e1.put_grid(0, T0.numpy())
        # ghosts of the whole box: images of its own atoms (positions follow from the unperturbed lattice's shifts)
        wl = H.make_system(cells, ntypes=a.elements, with_list=False)
        gshift = wl["x"][wl["nlocal"]:] - wl["x"][own]
        w["x"][w["nlocal"]:] = w["x"][own] + gshift
        w["v"][w["nlocal"]:] = w["v"][own]
        m1 = Verlet(e1, w, plan1, None, dev, a1, torch)
        m1.f[:] = f0[order]
        for k in range(nsteps):
            m1.step(k0 + 1 + k)
        inv = torch.empty_like(order)
        inv[order] = torch.arange(len(order), device=dev)
        def dev_rel(b, r):
            r = r[inv] if r.shape[0] == natoms else r
            return float((b - r).abs().max() / r.abs().max())
        rho1 = torch.as_tensor(e1.probe(0)[: w["nlocal"]], device=dev).reshape(-1, 1)
        T1 = e1.get_grid(0)
        res = {"steps": nsteps, "max_rel_dev_f": dev_rel(fb, m1.f), "max_rel_dev_rho": dev_rel(rb, rho1),
               "max_rel_dev_x": dev_rel(xb, m1.x[: w["nlocal"]]), "max_rel_dev_T_e": float(np.abs(Tb - T1).max() / np.abs(T1).max()),
               "mean_T_e": float(T1.mean()), "sum_abs_f": float(m1.f.abs().sum()),
               "note": "bricks on %d GPUs (engine's NCCL data plane) against the whole box on one GPU, all atoms and grid cells, "
                       "after %d trajectory steps from the same state; scaled by the largest magnitude" % (world, nsteps)}
        res["max_rel_dev"] = max(res["max_rel_dev_f"], res["max_rel_dev_rho"], res["max_rel_dev_x"], res["max_rel_dev_T_e"])
        e1.close()
    dist.barrier()
    return res


def run_e2e(a, torch, eng, md, s, plan, D, timed, natoms, world):
    """the same metric through the C ABI with pinned HOST buffers (what FixEPHB200 does with LAMMPS' host arrays):
    x, v, f up and f, E down every step, atoms registered and the list rebuilt on the device every 10 steps"""
    nl, ng = s["nlocal"], s["nghost"]
    nt = nl + ng
    pin = lambda t: t.cpu().pin_memory()
    h_x, h_v = pin(md.x), pin(md.v)
    h_f = torch.zeros((nl, 3), dtype=torch.float64).pin_memory()
    h_type, h_mask, h_tag, h_owner = pin(md.type), pin(md.mask), pin(md.tag), pin(md.owner)
    nx, nv, nf = h_x.numpy(), h_v.numpy(), h_f.numpy()

    def reneighbor():
        eng.set_atoms(nl, ng, h_type.numpy(), h_mask.numpy(), h_tag.numpy(), h_owner.numpy())
        eng.build_neighbors(nx, CUTOFF)                    # list built on the device from the uploaded positions
        if D:
            eng.set_ghost_map(plan)

    def step_e2e(k):
        if k % REBUILD_EVERY == 0:
            reneighbor()
        eng.post_force(nx, nv, nf, None, k)         # x, v, f up (f behind the density pass), ghost exchange, f down
        return eng.end_of_step(None, nv)            # v up, source all-reduce + solve, E_local down

    reneighbor()
    for k in range(1, min(a.warmup, 3) + 1):
        step_e2e(k)
    ms_dev, ms_wall = timed(step_e2e, a.steps, a.warmup + 1)
    ms_e = max(ms_dev, ms_wall)
    rebuilds = len([k for k in range(a.warmup + 1, a.warmup + a.steps + 1) if k % REBUILD_EVERY == 0])
    topo = nt * (4 + 4 + 8) + ng * 4
    list_bytes = (topo + nt * 24) * rebuilds / a.steps
    h2d = 2 * nt * 24 + nl * 24 + nl * 24 + list_bytes
    d2h = nl * 24 + 8
    # ---- the same through the device-resident integration mode (fix keyword `integrate device`): x, v, f stay on the
    # device between the hooks; per step the pair forces go up, x, f and v come down; all four hooks are inside ----
    resident = None
    if world == 1:
        mass = np.array([0.0] + [MASS] * s["ntypes"])
        dtv, dtf = a.dt, 0.5 * a.dt * FTM2V
        nxl, nvl = nx[:nl], nv[:nl]

        def reneighbor_res():
            eng.set_atoms(nl, ng, h_type.numpy(), h_mask.numpy(), h_tag.numpy(), h_owner.numpy())
            eng.resident_upload(nx, nv)                                  # x, v up once ...
            eng.build_neighbors(None, CUTOFF)                            # ... and the list is built from the resident copy

        def step_res(k):
            rebuild = k % REBUILD_EVERY == 0
            # x down; without a re-neighbouring ahead the density pass of this step starts while x travels
            eng.resident_initial_integrate(nf, mass, dtv, dtf, nxl, start_post_force_step=-1 if rebuild else k)
            if rebuild:
                eng.resident_get(1, out=nvl)                             # LAMMPS re-neighbours from its host arrays: v down too
                reneighbor_res()
            sync = k % REBUILD_EVERY == 0                               # `sync 10`: LAMMPS' own f and v every 10th step
            eng.resident_post_force(nf, None, k, f_out=nf if sync else None)    # pair forces up (total forces down)
            eng.resident_final_integrate(mass, dtf, nvl if sync else None)      # (v down)
            return eng.resident_end_of_step(True)                        # E_local down

        reneighbor_res()
        for k in range(1, min(a.warmup, 3) + 1):
            step_res(k)
        r_dev, r_wall = timed(step_res, a.steps, a.warmup + 1)
        ms_r = max(r_dev, r_wall)
        resident = {"value": natoms * a.steps / (ms_r * 1e-3), "unit": UNIT, "ms_per_step": ms_r / a.steps,
                    "h2d_bytes_per_step": int(nl * 24 + (topo + 2 * nt * 24 + nt * 24) * rebuilds / a.steps),
                    "d2h_bytes_per_step": int(nl * 24 + 8 + 3 * nl * 24 * rebuilds / a.steps),
                    "note": "device-resident integration (eph_b200_resident_*; FixEPHB200 keywords `integrate device sync %d`): all "
                            "four hooks, x / v / f stay on the device; per step the pair forces go up and x comes down; every %d "
                            "steps f and v come down too (thermo / dump cadence), v once more for the re-neighbouring, atoms are "
                            "registered, the list is rebuilt on the device and x, v are uploaded" % (REBUILD_EVERY, REBUILD_EVERY)}
    host_mode = {"value": natoms * a.steps / (ms_e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(h2d * world),
                 "d2h_bytes_per_step": int(d2h * world), "ms_per_step": ms_e / a.steps,
            "note": "pinned host x, v, f through the C ABI (HOST memspace): x, v, f up and f, E down every step; every %d steps "
                    "the atom arrays are re-registered and the neighbour list is rebuilt on the device from the positions "
                    "(eph_b200_build_neighbors); post_force + end_of_step (the integrator hooks of FixEPHB200 are host loops "
                    "in this mode); bytes summed over ranks" % REBUILD_EVERY}
    if resident is None:
        return host_mode
    # one GPU: the headline is the mode that keeps x, v, f on the device (all four hooks); the plain host mode next to it
    return dict(resident, mode="integrate device (resident x, v, f)", host_mode=host_mode)


def c4_leg(a, torch, lib, host, P, s, gridn, box, local, stream, dev, natoms, steps=10):
    """config C4 on the same positions and list: four elements (types uniform random), the NiCoCrFe table: two table
    look-ups per pair and a second pair-weight stream"""
    a4 = argparse.Namespace(**vars(a))
    a4.elements = 4
    s4 = dict(s)
    s4["ntypes"] = 4
    s4["type"] = np.random.default_rng(7).integers(1, 5, len(s["type"])).astype(np.int32)
    s4["type"][s["nlocal"]:] = s4["type"][s["ghost_owner"]]
    eng = make_engine(lib, host, a4, s4, gridn, box, local, 0, 1, stream, elements=4)
    md = Verlet(eng, s4, P.ExchangePlan(s4, 0, 1), None, dev, a4, torch)
    for k in range(1, 4):
        md.step(k)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    eng.set_profiling(True)
    e0.record()
    for k in range(4, 4 + steps):
        md.step(k)
    e1.record()
    torch.cuda.synchronize()
    kt = eng.kernel_times()
    ms = e0.elapsed_time(e1) / steps
    eng.close()
    return {"value": natoms / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "steps": steps,
            "kernels_ms": {k: round(v[0] / max(v[1], 1), 4) for k, v in kt.items()},
            "note": "same box and list, 4 elements (NiCoCrFe_PRB2019), %s mode, one re-neighbouring inside" % a.mode}


def small_box_check(a, torch, lib, host, local, stream, dev, cells=16, nsteps=3):
    """a 16 384-atom box with the PKA through the same engine build, against the unmodified reference fix on the host
    (oracle/_ref; the oracle port if it is not there): all atoms, forces / densities / grid"""
    from eph_harness import harness as H
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import traj
    s = H.make_system(cells)
    s["v"][0] = V_PKA * np.array([0.835115, 0.543981, 0.081652])
    s["v"][s["nlocal"]:][s["ghost_owner"] == 0] = s["v"][0]
    box = [0, s["box"][0], 0, s["box"][1], 0, s["box"][2]]
    xis = [np.random.default_rng(40 + k).normal(size=(s["nlocal"], 3)) for k in range(nsteps)]
    eng = lib.Engine([0], flags=7, device=local, stream=stream)
    eng.set_tables_from(host.BetaTables(path=beta_file(1)))
    eng.set_grid(8, 8, 8, box, 300.0, 1.0, 3.5e-6, 0.1248)
    eng.set_dt(a.dt)
    eng.set_skin(2.0)
    t = lambda arr, dt: torch.as_tensor(np.ascontiguousarray(arr), dtype=dt, device=dev)
    eng.set_atoms(s["nlocal"], s["nghost"], t(s["type"], torch.int32), t(s["mask"], torch.int32), t(s["tag"], torch.int64),
                  t(s["ghost_owner"], torch.int32))
    eng.build_neighbors(t(s["x"], torch.float64), CUTOFF)      # the list the bench runs on: built by the engine
    recs = traj.run_engine(eng, s, xis, [MASS], a.dt, device=True)
    from oracle import reference as R
    if R.available():
        devnull, saved = os.open(os.devnull, os.O_WRONLY), os.dup(1)
        os.dup2(devnull, 1)
        try:
            drv = R.fix_driver(s, H.fix_args(7, beta_file(1), ["Ni"], grid=(8, 8, 8)), dt=a.dt)
            refs = traj.run_fix_driver(drv, s, xis)
        finally:
            os.dup2(saved, 1)
        kind = "reference (oracle/_ref)"
    else:
        from oracle import oracle as O
        fx = O.Fix(s, O.Beta(path=beta_file(1)), O.FDM(8, 8, 8, box, 300.0, 3.5e-6, 1.0, 0.1248), 7, dt=a.dt)
        refs = traj.run_oracle(fx, s, xis, [MASS])
        kind = "oracle port"
    worst = {}
    for ra, rb in zip(recs, refs):
        for key in ("f", "array", "T", "x", "v"):
            worst[key] = max(worst.get(key, 0.0), H.error_metrics(ra[key], rb[key]))
    eng.close()
    return {"atoms": s["nlocal"], "steps": nsteps, "against": kind, "max_rel_dev": max(worst.values()), "by_quantity": worst,
            "records": "packed from step 2 on (step 1 builds the inner list on the fp64 records)"}


def fdm_bench(lib, host, stream, local, peak, n=256, solves=5):
    """FDM Mcell-updates/s on a 256^3 grid with 13 sub-steps per solve (TB_Bench fine-grid parameters): the
    constant-coefficient path (the grid `fix eph` creates without a grid file) and the general variable-coefficient path"""
    import torch
    L = 56.32 * n / 128.0
    out = {}
    for kind in ("uniform", "general"):
        eng = lib.Engine([0], flags=7, device=local, stream=stream)
        eng.set_tables_from(host.BetaTables(path=beta_file(1)))
        kap = 0.01248 if kind == "uniform" else 0.01248 * (0.9 + 0.1 * np.random.default_rng(1).random(n ** 3))
        eng.set_grid(n, n, n, [0, L, 0, L, 0, L], 300.0, 1.0, 3.5e-6, kap)    # TB_Bench 03_Grid_Fine parameters
        eng.set_dt(1e-4)
        x = np.array([[1.0, 1.0, 1.0]]); z = np.zeros((1, 3))
        eng.set_atoms(1, 0, np.array([1], dtype=np.int32), np.array([0], dtype=np.int32), np.array([1], dtype=np.int64))
        eng.set_neighbors(np.array([0, 0], dtype=np.int64), np.array([0], dtype=np.int32))
        dx, dv, df = (torch.as_tensor(t, device=torch.device("cuda", local)) for t in (x, z, z.copy()))
        eng.post_force(dx, dv, df, None, 0)
        for _ in range(2):
            eng.end_of_step(dx, dv, want_energy=False)
        sub = eng.last_substeps()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        eng.set_profiling(True)
        e0.record()
        for _ in range(solves):
            eng.end_of_step(dx, dv, want_energy=False)
        e1.record()
        torch.cuda.synchronize()
        kt = eng.kernel_times().get("fdm_substep", (0.0, 1))
        ms = e0.elapsed_time(e1)
        cells = n ** 3
        k_ms = kt[0] / max(kt[1], 1)
        eng.close()
        # the constant-coefficient kernel moves 24 B per cell-update (T_e in/out, dT_e in), the general one SURVEY's 60 B
        nbytes = 24.0 if kind == "uniform" else 60.0
        gbs = nbytes * cells / (k_ms * 1e-3) / 1e9
        out[kind] = {"value": cells * sub * solves / (ms * 1e-3) / 1e6, "unit": "Mcell-updates/s", "substeps": sub, "solves": solves,
                     "ms_per_solve": ms / solves,
                     "roofline": {"bound": "hbm", "kernel": "fdm_substep (%s TMA path)" % ("constant-coefficient" if kind == "uniform" else "general"),
                                  "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak,
                                  "algorithmic_bytes_per_cell_update": nbytes, "kernel_ms": k_ms}}
    return {"metric": "FDM Mcell-updates/s", "value": out["general"]["value"], "unit": "Mcell-updates/s", "grid": [n] * 3,
            "general_path": out["general"], "constant_coefficient_path": out["uniform"],
            "note": "value = the general variable-coefficient path (60 B per cell-update, SURVEY 8d); the constant-coefficient "
                    "path is what `fix eph ... NX NY NZ NULL` runs"}


def main():
    a = parse_args()
    if a.impl == "reference":
        run_reference_arm(a)
    else:
        run_b200(a)


if __name__ == "__main__":
    main()
