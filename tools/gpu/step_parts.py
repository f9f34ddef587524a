#!/usr/bin/env python
"""Where the trajectory step of bench.py spends its time, part by part (development probe, not a bench value).

    python tools/gpu/step_parts.py --cells 50                 # one GPU, the brick size of an 8-GPU run
    torchrun --nproc-per-node 2 ... tools/gpu/step_parts.py   # bricks over NCCL

Every part of Verlet.step is bracketed by synchronize (+ barrier on several ranks) and timed on the wall clock, `reps`
times; then the unbracketed loop for comparison.  Prints one JSON line on rank 0.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench as B  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cells", type=int, default=100)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--overlap", action="store_true")
    ap.add_argument("--neigh", default="lammps")
    b = ap.parse_args()
    import torch
    import torch.distributed as dist
    from eph_b200 import host, lib
    from eph_harness import parallel as P

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    sys.argv = [sys.argv[0], "--gpus", str(world), "--cells", str(b.cells), "--neigh", b.neigh] + (["--overlap"] if b.overlap else [])
    a = B.parse_args()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    D = dist if world > 1 else None
    grid = P.brick_grid(world)
    s = B.build_workload(a.cells, brick=(rank, grid) if world > 1 else None)
    if world == 1:
        s["grid"] = (1, 1, 1)
    box = [0, s["box"][0], 0, s["box"][1], 0, s["box"][2]]
    plan = P.ExchangePlan(s, rank, world, D)
    tstream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(tstream)
    eng = B.make_engine(lib, host, a, s, (a.grid,) * 3, box, local, rank, world, tstream.cuda_stream)
    if D:
        P.attach_comm(eng, D, rank, world)
        gstream = torch.cuda.Stream(device=dev)
        eng.set_grid_stream(gstream.cuda_stream)
        if a.overlap:
            cstream = torch.cuda.Stream(device=dev)
            eng.set_comm_stream(cstream.cuda_stream)
    md = B.Verlet(eng, s, plan, D, dev, a, torch)
    for k in range(1, 6):
        md.step(k)

    def sync():
        torch.cuda.synchronize()
        if D:
            D.barrier()
            torch.cuda.synchronize()

    parts = {}

    def timed(name, fn):
        sync()
        t0 = time.perf_counter()
        fn()
        t_host = time.perf_counter() - t0
        sync()
        dt = time.perf_counter() - t0
        p = parts.setdefault(name, [0.0, 0.0, 0])
        p[0] += dt; p[1] += t_host; p[2] += 1

    k = 100
    for r in range(b.reps):
        k += 1
        timed("empty", lambda: None)
        timed("initial_integrate", lambda: eng.initial_integrate(md.x, md.v, md.f, md.mass, md.dtv, md.dtf))
        timed("refresh_ghosts", md.refresh_ghosts)
        if r % 4 == 0:
            timed("register.set_atoms", lambda: eng.set_atoms(md.nl, md.ng, md.type, md.mask, md.tag, md.owner))
            if md.csr is not None:
                timed("register.set_neighbors", lambda: eng.set_neighbors(*md.csr))
            else:
                timed("register.build_neighbors", lambda: eng.build_neighbors(md.x, B.CUTOFF))
            if world > 1:
                timed("register.set_ghost_map", lambda: eng.set_ghost_map(plan))
            timed("register.refresh_ghosts", md.refresh_ghosts)
            timed("force_clear", lambda: md.f.zero_())
            timed("post_force(after register)", lambda: eng.post_force(md.x, md.v, md.f, None, k))
        else:
            timed("force_clear", lambda: md.f.zero_())
            timed("post_force", lambda: eng.post_force(md.x, md.v, md.f, None, k))
        timed("final_integrate", lambda: eng.final_integrate(md.v, md.f, md.mass, md.dtf))
        timed("end_of_step", lambda: eng.end_of_step(None, md.v, want_energy=False))
    # the loop as bench.py runs it
    sync()
    t0 = time.perf_counter()
    n = 40
    for i in range(n):
        md.step(1000 + i)
    t_host = time.perf_counter() - t0
    sync()
    loop = (time.perf_counter() - t0) / n
    out = {"world": world, "cells": b.cells, "nlocal": md.nl, "nghost": md.ng, "overlap": b.overlap,
           "loop_ms_per_step": 1e3 * loop, "loop_host_ms_per_step": 1e3 * t_host / n,
           "parts_ms": {kk: {"total": round(1e3 * v[0] / v[2], 4), "host_enqueue": round(1e3 * v[1] / v[2], 4), "n": v[2]} for kk, v in parts.items()}}
    if rank == 0:
        print(json.dumps(out), flush=True)
    if D:
        D.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
