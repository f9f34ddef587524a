"""Host-side logic of the multi-GPU path on CPU: brick decomposition and the ghost exchange plan, with two
gloo ranks.  (The compute itself needs GPUs: tests/test_gpu_parity.py::test_two_rank_engine_matches_oracle.)"""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from eph_b200 import harness as H
from eph_b200 import parallel as P


def test_brick_grid_shapes():
    assert P.brick_grid(1) == (1, 1, 1) and P.brick_grid(2) == (2, 1, 1)
    assert P.brick_grid(4) == (2, 2, 1) and P.brick_grid(8) == (2, 2, 2)


def test_bricks_partition_the_box_and_ghost_shells_are_complete():
    grid = (2, 2, 1)
    whole = H.make_system(6)
    seen = []
    for r in range(4):
        s = H.make_system(6, brick=(r, grid))
        seen.append(s["tag"][: s["nlocal"]])
        # every list entry within the cut-off is present: compare one atom's in-cutoff neighbour tags with the whole box
        i = 3
        t = s["tag"][i]
        nb = np.sort(s["tag"][s["neigh"][s["offsets"][i]: s["offsets"][i + 1]]])
        iw = int(np.nonzero(whole["tag"][: whole["nlocal"]] == t)[0][0])
        nbw = np.sort(whole["tag"][whole["neigh"][whole["offsets"][iw]: whole["offsets"][iw + 1]]])
        assert np.array_equal(nb, nbw)
        assert np.array_equal(P.owner_rank_of(s["x"][: s["nlocal"]], s["box"], grid), np.full(s["nlocal"], r))
    alltags = np.sort(np.concatenate(seen))
    assert np.array_equal(alltags, np.arange(1, whole["natoms"] + 1))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q, cells=6):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        grid = P.brick_grid(world)
        s = H.make_system(cells, brick=(rank, grid))
        plan = P.ExchangePlan(s, rank, world, dist)
        nl, ng = s["nlocal"], s["nghost"]
        # payload = a function of the atom tag; after the exchange every ghost must hold its owner's payload
        pay = np.zeros((nl + ng, 4))
        pay[:nl] = np.stack([s["tag"][:nl] * 1.0, s["tag"][:nl] * 2.0, -s["tag"][:nl] * 1.0, np.sqrt(s["tag"][:nl])], axis=1)
        send = torch.as_tensor(pay[plan.flat_send_index()].reshape(-1))
        recv = torch.empty(4 * sum(plan.recv_counts), dtype=torch.float64)
        dist.all_to_all_single(recv, send, output_split_sizes=[4 * c for c in plan.recv_counts],
                               input_split_sizes=[4 * c for c in plan.send_counts])
        pay[plan.flat_recv_index()] = recv.numpy().reshape(-1, 4)
        own = plan.self_owner >= 0
        pay[nl:][own] = pay[plan.self_owner[own]]
        gt = s["tag"][nl:].astype(np.float64)
        ok = np.array_equal(pay[nl:], np.stack([gt, 2 * gt, -gt, np.sqrt(gt)], axis=1))
        # grid all-reduce: each rank deposits its own atoms, the sum must equal the whole-box deposit
        cell = np.minimum((s["x"][:nl] / (s["box"] / 4)).astype(int), 3)
        hist = np.zeros(64)
        np.add.at(hist, cell[:, 0] + 4 * cell[:, 1] + 16 * cell[:, 2], 1.0)
        t = torch.as_tensor(hist)
        dist.all_reduce(t)
        q.put((rank, bool(ok), float(t.sum()), int((~own).sum()), int(own.sum())))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("cells", [6, (10, 5, 5)])   # cubic box (strong scaling); 5^3 cells per rank (bench.py --weak)
def test_exchange_plan_two_gloo_ranks(cells):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q, cells)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(60)
    natoms = 4 * (cells ** 3 if np.isscalar(cells) else int(np.prod(cells)))
    for rank, ok, total, remote, own in res:
        assert ok, "ghost payload mismatch on rank %d" % rank
        assert total == natoms
        assert remote > 0 and own > 0     # 2 x 1 x 1 bricks: x-ghosts are remote, y/z images are the rank's own atoms


# ---- sharded grid solve: the orchestration of eph_b200.parallel.sharded_grid_solve (slabs, halo planes, all-gather) ----
class NumpySlabEngine:
    """Stands in for eph_b200.lib.Engine (same four members the orchestration uses) with a numpy form of the
    reference's explicit sub-step (eph_fdm.h:319-395, all cells dynamic).  Planes outside the slab are poisoned with
    NaN in the buffer a sub-step writes, so a missing or misplaced halo plane cannot go unnoticed."""

    def __init__(self, shape, box, c, dt, src):
        self.grid_shape = shape
        nx, ny, nz = shape
        r3 = lambda a: np.asarray(a, dtype=np.float64).reshape(nz, ny, nx).copy()
        self.T = [torch.as_tensor(r3(c["T"]).reshape(-1)), torch.full((nx * ny * nz,), float("nan"), dtype=torch.float64)]
        self.cur = 0
        self.K, self.C, self.rho, self.S, self.src = r3(c["kap"]), r3(c["Ce"]), r3(c["rho"]), r3(c["S"]), r3(src)
        d = [(box[1] - box[0]) / nx, (box[3] - box[2]) / ny, (box[5] - box[4]) / nz]
        self.inv = [1.0 / (q * q) for q in d]
        # sub-step count of the reference (eph_fdm.h:290-313), steps = 1
        r = dt * sum(self.inv) / self.C.min() / self.rho.min() * self.K.max()
        self.n = 1
        if r > 0.4:
            self.n = max(int(dt / (0.4 * dt / r)), 1)
        self.inner_dt = dt / self.n

    def grid_plan_substeps(self):
        return self.n

    def grid_tensor(self, which=0):
        return self.T[self.cur]

    def grid_substep(self, z0, z1):
        nx, ny, nz = self.grid_shape
        Tin = self.T[self.cur].numpy().reshape(nz, ny, nx)
        Tout = self.T[1 - self.cur].numpy().reshape(nz, ny, nx)
        Tout[...] = np.nan
        for k in range(z0, z1):
            T, kr = Tin[k], self.K[k]
            dd = np.zeros_like(T)
            for axis, inv in ((1, self.inv[0]), (0, self.inv[1])):
                Tq, Tp = np.roll(T, -1, axis), np.roll(T, 1, axis)
                dd += (np.roll(kr, -1, axis) - np.roll(kr, 1, axis)) * (Tq - Tp) * inv * 0.25
                dd += kr * ((Tq + Tp - 2.0 * T) * inv)
            Tq, Tp = Tin[(k + 1) % nz], Tin[(k - 1) % nz]
            dd += (self.K[(k + 1) % nz] - self.K[(k - 1) % nz]) * (Tq - Tp) * self.inv[2] * 0.25
            dd += kr * ((Tq + Tp - 2.0 * T) * self.inv[2])
            Tout[k] = np.maximum(T + (dd + self.src[k] + self.S[k]) / (self.rho[k] * self.C[k]) * self.inner_dt, 0.0)
        self.cur = 1 - self.cur


def _slab_case(shape, seed=71):
    rng = np.random.default_rng(seed)
    n = int(np.prod(shape))
    c = dict(T=300 + 100 * rng.random(n), kap=0.1248 * (0.5 + rng.random(n)), Ce=3.5e-6 * (0.5 + rng.random(n)),
             S=1e-3 * rng.random(n), rho=1.0 + 0.2 * rng.random(n))
    return c, 1e-2 * rng.normal(size=n)


SLAB_BOX = [0.0, 17.6, -1.0, 16.6, 2.0, 19.6]


def _slab_worker(rank, world, port, q, shape, dt):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        c, src = _slab_case(shape)
        eng = NumpySlabEngine(shape, SLAB_BOX, c, dt, src)
        n = P.sharded_grid_solve(eng, dist, rank, world)
        q.put((rank, n, eng.grid_tensor(0).numpy().copy()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,shape,dt", [(2, (6, 5, 4), 5e-3), (2, (5, 3, 2), 5e-3), (3, (4, 4, 6), 2e-3), (2, (6, 5, 4), 1e-4)])
def test_sharded_grid_solve_matches_the_whole_grid_solve(world, shape, dt):
    """slab sub-steps + halo planes + all-gather over gloo ranks == one rank solving the whole grid == the oracle's
    EPH_FDM::solve; two ranks exercise the prev == next pairing of the halo exchange, one-plane slabs the wrap"""
    from oracle import oracle as O
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_slab_worker, args=(r, world, port, q, shape, dt)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=180) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(60)
    c, src = _slab_case(shape)
    one = NumpySlabEngine(shape, SLAB_BOX, c, dt, src)
    assert P.sharded_grid_solve(one, None, 0, 1) == one.n
    whole = one.grid_tensor(0).numpy()
    o = O.FDM(*shape, SLAB_BOX, 300.0, 3.5e-6, 1.0, 0.1248)
    for which, key in ((0, "T"), (1, "S"), (2, "rho"), (3, "Ce"), (4, "kap")):
        o.field(which)[:] = c[key]
    o.field(5)[:] = src
    o.set_dt(dt)
    o.solve()
    assert H.error_metrics(whole, o.field(0)) < 1e-12
    for rank, n, T in res:
        assert n == one.n and (n > 1 or dt < 1e-3)
        assert np.all(np.isfinite(T)), "rank %d read a plane nobody sent" % rank
        assert np.array_equal(T, whole), "rank %d" % rank


def test_grid_slabs_partition_the_planes():
    assert P.grid_slab(64, 3, 8) == (24, 32) and P.grid_slab(6, 0, 1) == (0, 6)
    assert P.grid_slab(10, 0, 4) is None
    for world in (1, 2, 4, 8):
        planes = [z for r in range(world) for z in range(*P.grid_slab(64, r, world))]
        assert planes == list(range(64))
