"""GPU tests of paths added after this round's GPU budget was spent (their first run on a B200 is the driver's
round-end run).  They live in the last test file so that `-x` reaches them only after the suite measured earlier in
the round has passed."""
import os

import numpy as np
import pytest

from eph_harness import harness as H
from eph_b200 import host, lib
from eph_harness import parallel as P
from oracle import oracle as O

import traj
from conftest import GOLDEN

pytestmark = pytest.mark.gpu
TOL = 1e-10


def test_fix_b200_peratom_cadence():
    """keyword `peratom N`: array_atom is refreshed from the device in steps divisible by N only (forces are unaffected)"""
    g = np.load(os.path.join(GOLDEN, "caseA_example1.npz"))
    s = traj.system_from_golden(g)
    s["natoms"] = s["nlocal"]
    cwd = os.getcwd()
    os.chdir(GOLDEN)
    try:
        drv = host.FixDriver(s, H.fix_args(3, "Ni_trunc.beta", ["Ni"], grid=(1, 1, 1), style="eph/b200",
                                           extra=["rng", "mars", "peratom", "2"]), dt=float(g["dt"]))
    finally:
        os.chdir(cwd)
    recs = traj.run_fix_driver(drv, s, list(g["xi"]))
    for k in range(3):
        assert H.error_metrics(recs[k]["f"], g["out_f"][k]) < TOL
    assert not recs[0]["array"].any()
    assert H.error_metrics(recs[1]["array"], g["out_array"][1]) < TOL
    assert np.array_equal(recs[2]["array"], recs[1]["array"])


def _dummy_atom(eng):
    """a one-atom system outside the fix group: end_of_step then only touches the grid"""
    x = np.array([[1.0, 1.0, 3.0]]); z = np.zeros((1, 3))
    eng.set_atoms(1, 0, np.array([1], dtype=np.int32), np.array([0], dtype=np.int32), np.array([1], dtype=np.int64))
    eng.set_neighbors(np.array([0, 0], dtype=np.int64), np.array([0], dtype=np.int32))
    return x, z


def _views(engs, nz, plane):
    """the engines' current T_e buffers as [nz][plane] torch views of device memory (tests/test_engine_emulated.py
    swaps in numpy views of the host build's memory)"""
    import torch
    for e in engs:
        e.synchronize()
    out = [e.grid_tensor(0).view(nz, plane) for e in engs]
    torch.cuda.synchronize()
    return out


def _copy(dst, src):
    if hasattr(dst, "copy_"):
        dst.copy_(src)
    else:
        dst[...] = src


def _solve_sharded(engs, xz, shape):
    """what eph_harness.parallel.sharded_grid_solve does over NCCL, with the ranks' engines in one process: slab sub-steps,
    halo planes copied between the engines' T_e arrays, slabs gathered at the end"""
    W = len(engs)
    nx, ny, nz = shape
    plane = nx * ny
    slabs = [P.grid_slab(nz, r, W) for r in range(W)]
    for e in engs:
        x, z = xz
        e.post_force(x, z, z.copy(), None, 0)
        e.end_of_step_begin(x, z)
    ns = [e.grid_plan_substeps() for e in engs]
    assert len(set(ns)) == 1
    n = ns[0]

    def views():
        return _views(engs, nz, plane)

    def sync():
        try:
            import torch
            if torch.cuda.is_available():
                torch.cuda.synchronize()
        except ImportError:
            pass

    for s in range(n):
        for e, (z0, z1) in zip(engs, slabs):
            e.grid_substep(z0, z1)
        if s < n - 1 and W > 1:
            Ts = views()
            for r, (z0, z1) in enumerate(slabs):
                zlo, zhi = (z0 - 1) % nz, z1 % nz
                _copy(Ts[r][zlo], Ts[(r - 1) % W][zlo])
                _copy(Ts[r][zhi], Ts[(r + 1) % W][zhi])
            sync()
    if W > 1 and n > 0:
        Ts = views()
        for r in range(W):
            for q, (z0, z1) in enumerate(slabs):
                if q != r:
                    _copy(Ts[r][z0:z1], Ts[q][z0:z1])
        sync()
    for e in engs:
        e.end_of_step_end(True, external=True)
    return n


@pytest.mark.parametrize("shape,world,kind", [((32, 6, 8), 2, "walls"), ((32, 6, 8), 4, "walls"), ((33, 9, 6), 3, "walls"),
                                              ((32, 16, 8), 2, "uniform"), ((64, 8, 16), 8, "uniform"), ((16, 4, 4), 4, "general")])
def test_sharded_grid_solve_matches_replicated_solve_and_oracle(synth_beta_1, shape, world, kind):
    """The z-range of the three sub-step kernels (TMA general, TMA constant-coefficient, plain): `world` engines each
    advance their slab with halo planes exchanged between sub-steps; every engine must end with exactly the field
    the replicated solve produces, and that field matches the oracle."""
    rng = np.random.default_rng(47)
    box = [0.0, 35.2, 0.0, 17.6, -3.0, 14.6]
    n = int(np.prod(shape))
    o = O.FDM(*shape, box, 300.0, 3.5e-6, 1.0, 0.1248)
    engs = [lib.Engine([0], flags=7) for _ in range(world + 1)]   # the last one solves the whole grid itself
    T0 = 300 + 200 * rng.random(n)
    if kind == "uniform":
        for e in engs:
            e.set_tables_from(host.BetaTables(path=synth_beta_1))
            e.set_grid(*shape, box, 300.0, 1.0, 3.5e-6, 0.1248)
            e.put_grid(0, T0)
        o.field(0)[:] = T0
    else:
        fl = np.ones(n, dtype=np.int16)
        if kind == "walls":
            fl[rng.random(n) < 0.15] = 2
            fl[rng.random(n) < 0.1] = 0
        c = dict(T=T0, kap=0.1248 * (0.5 + rng.random(n)), Ce=3.5e-6 * (0.5 + rng.random(n)), S=1e-3 * rng.random(n),
                 rho=1.0 + 0.2 * rng.random(n))
        for which, key in ((0, "T"), (1, "S"), (2, "rho"), (3, "Ce"), (4, "kap")):
            o.field(which)[:] = c[key]
        o.flags()[0][:] = fl
        for e in engs:
            e.set_tables_from(host.BetaTables(path=synth_beta_1))
            e.set_grid(*shape, box, c["T"], c["rho"], c["Ce"], c["kap"], S_e=c["S"], flag=fl)
    xz = None
    for e in engs:
        xz = _dummy_atom(e)
    for dt in (1e-4, 2e-3):   # the second one needs sub-steps (r > 0.4)
        o.set_dt(dt)
        for e in engs:
            e.set_dt(dt)
        for _ in range(2):
            src = 1e-2 * rng.normal(size=n)
            o.field(5)[:] = src
            for e in engs:
                e.put_grid(5, src)          # every rank holds the all-reduced source term
            o.solve()
            x, z = xz
            engs[-1].post_force(x, z, z.copy(), None, 0)
            engs[-1].end_of_step(x, z)
            nsub = _solve_sharded(engs[:-1], xz, shape)
            assert nsub == engs[-1].last_substeps()
            whole = engs[-1].get_grid(0)
            assert H.error_metrics(whole, o.field(0)) < TOL
            for r, e in enumerate(engs[:-1]):
                assert np.array_equal(e.get_grid(0), whole), "rank %d of %d" % (r, world)
                assert np.all(e.get_grid(5) == 0.0)
    assert engs[0].last_substeps() > 1


def test_fix_integrate_device_matches_reference():
    """keyword `integrate device` of FixEPHB200: the velocity-Verlet half steps run on the device, x, v, f stay there between
    the hooks; against the compiled reference fix through re-neighbourings in LAMMPS' order (between initial_integrate and
    post_force)"""
    import reneighbour_cases
    reneighbour_cases.resident_case(cells=6, steps=8, every=3)
    reneighbour_cases.resident_case(cells=6, steps=8, every=3, rng="philox", sync=2)
