"""GPU tests of paths added after this round's GPU budget was spent (their first run on a B200 is the driver's
round-end run).  They live in the last test file so that `-x` reaches them only after the suite measured earlier in
the round has passed."""
import os

import numpy as np
import pytest

from eph_b200 import harness as H
from eph_b200 import host, lib
from eph_b200 import parallel as P
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


def _solve_sharded(engs, xz, shape):
    """what eph_b200.parallel.sharded_grid_solve does over NCCL, with the ranks' engines in one process: slab sub-steps,
    halo planes copied between the engines' T_e arrays, slabs gathered at the end"""
    import torch
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
        for e in engs:
            e.synchronize()
        return [e.grid_tensor(0).view(nz, plane) for e in engs]

    for s in range(n):
        for e, (z0, z1) in zip(engs, slabs):
            e.grid_substep(z0, z1)
        if s < n - 1 and W > 1:
            Ts = views()
            for r, (z0, z1) in enumerate(slabs):
                zlo, zhi = (z0 - 1) % nz, z1 % nz
                Ts[r][zlo].copy_(Ts[(r - 1) % W][zlo])
                Ts[r][zhi].copy_(Ts[(r + 1) % W][zhi])
            torch.cuda.synchronize()
    if W > 1 and n > 0:
        Ts = views()
        for r in range(W):
            for q, (z0, z1) in enumerate(slabs):
                if q != r:
                    Ts[r][z0:z1].copy_(Ts[q][z0:z1])
        torch.cuda.synchronize()
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


# ---------------------------------------------------------------------------------------------------------------------
# `fix eph/atomic` on the device (csrc/eph_atomic.cu behind include/eph_b200_atomic.h).  The same
# cases run on the CPU against a host build of the same source (tests/test_atomic_emulated.py); these are the real
# thing: the sm_100a kernels with 8 lanes per atom, through the C ABI and through FixEPHAtomicB200.
# ---------------------------------------------------------------------------------------------------------------------
import atomic_cases as cases  # noqa: E402
from eph_b200 import atomic as A  # noqa: E402


@pytest.fixture(scope="module")
def make_engine():
    return lambda tb, tk, flags, **kw: A.AtomicEngine(tb, tk, flags, **kw)


@pytest.fixture(scope="module")
def kappa_tables():
    return A.KappaTables(cases.KAPPA)


@pytest.mark.parametrize("flags,loops,group_fraction", [(7, 0, None), (7, 3, None), (1, 0, None), (2, 0, None), (5, 2, None),
                                                        (6, 1, None), (7 + 16, 2, None), (7 + 32, 2, None), (7 + 8, 1, None),
                                                        (7, 2, 0.7), (4, 2, 0.5)])
def test_atomic_engine_matches_oracle(make_engine, kappa_tables, flags, loops, group_fraction):
    cases.trajectory_case(make_engine, kappa_tables, flags, loops, group_fraction)


def test_atomic_engine_larger_box(make_engine, kappa_tables):
    """4000 atoms: more CTAs than one, rows longer than one sweep of the 8 lanes"""
    cases.trajectory_case(make_engine, kappa_tables, 7, 2, None, n=10, steps=2)


def test_atomic_engine_two_elements(make_engine, kappa_tables, tmp_path):
    beta2 = str(H.write_beta_file(tmp_path / "synth2.beta", H.synthetic_knots(2, n_beta=5001, drho=0.01)))
    cases.trajectory_case(make_engine, kappa_tables, 7, 2, None, ntypes=2, beta=beta2, names=("Ni", "Co"))


def test_atomic_engine_heat_diffusion_from_gradient(make_engine, kappa_tables):
    cases.gradient_case(make_engine, kappa_tables)


@pytest.mark.parametrize("name", ["atomic_caseA", "atomic_caseB_group"])
def test_atomic_engine_matches_committed_golden_vectors(make_engine, kappa_tables, name):
    cases.golden_engine_case(make_engine, kappa_tables, name)


@pytest.mark.parametrize("name", ["atomic_caseA", "atomic_caseB_group"])
def test_fix_atomic_b200_matches_committed_golden_vectors(name):
    cases.golden_fix_case(lambda s, args: A.fix_driver(s, args), name)


def test_atomic_engine_builtin_gaussian_stream(make_engine, kappa_tables):
    cases.philox_case(make_engine, kappa_tables)


# ---------------------------------------------------------------------------------------------------------------------
# `fix eph/coloured/exp` on the device: model 4 + the exponential memory kernel (eph_b200_set_colour, colour_filter_kernel)
# ---------------------------------------------------------------------------------------------------------------------
def _coloured_engine(s, beta, flags, gb, tau0, dt=1e-4):
    box = [0, s["box"][0], 0, s["box"][1], 0, s["box"][2]]
    eng = lib.Engine([0], flags=flags, groupbit=gb)
    eng.set_tables_from(host.BetaTables(path=beta))
    eng.set_grid(2, 2, 2, box, 300.0, 1.0, 3.5e-6, 0.1248)
    eng.set_dt(dt)
    eng.set_colour(tau0)
    eng.set_atoms(s["nlocal"], s["nghost"], np.ascontiguousarray(s["type"], dtype=np.int32),
                  np.ascontiguousarray(s["mask"], dtype=np.int32), np.ascontiguousarray(s["tag"], dtype=np.int64),
                  np.ascontiguousarray(s["ghost_owner"], dtype=np.int32))
    eng.set_neighbors(np.ascontiguousarray(s["offsets"], dtype=np.int64), np.ascontiguousarray(s["neigh"], dtype=np.int32))
    return eng, box


@pytest.mark.parametrize("flags,group_fraction", [(7, None), (3, None), (1, None), (2, None), (7 + 16, None), (7 + 32, None),
                                                  (7 + 16 + 32, None), (7, 0.6)])
def test_coloured_engine_matches_oracle(ni_trunc_beta, flags, group_fraction):
    s = H.make_system(3, group_fraction=group_fraction)
    gb = 2 if group_fraction else 1
    tau0 = 5e-4
    eng, box = _coloured_engine(s, ni_trunc_beta, flags, gb, tau0)
    fx = O.Fix(s, O.Beta(path=ni_trunc_beta), O.FDM(2, 2, 2, box, 300.0, 3.5e-6, 1.0, 0.1248), flags, groupbit=gb, dt=1e-4)
    fx.set_colour(tau0)
    xis = [np.random.default_rng(i).normal(size=(s["nlocal"], 3)) if flags & 2 else None for i in range(4)]
    recs = traj.run_engine(eng, s, xis, [58.71], 1e-4, coloured=True)
    refs = traj.run_oracle(fx, s, xis, [58.71])
    for step, (a, b) in enumerate(zip(recs, refs)):
        for k in ("f", "array", "T", "w", "f_eph", "f_rng", "f_dis", "f_sto", "x", "v"):
            assert H.error_metrics(a[k], b[k]) < TOL, (step, k)
        assert abs(a["Ee"] - b["Ee"]) <= TOL * max(abs(b["Ee"]), 1e-300), step
    # a time-step change refreshes zeta (fix_eph_coloured_exp.cpp:686); the state carries over
    eng.set_dt(2e-4)
    fx.set_dt(2e-4)
    a = traj.run_engine(eng, s, xis[:1], [58.71], 2e-4, coloured=True)[0]
    # run_engine restarts from the system's initial x, v: do the same on the oracle side
    fx.x[...] = s["x"]; fx.v[...] = s["v"]
    b = traj.run_oracle(fx, s, xis[:1], [58.71])[0]
    for k in ("f", "f_dis", "f_sto"):
        assert H.error_metrics(a[k], b[k]) < TOL, k


def test_coloured_engine_device_pointers_and_state_round_trip(ni_trunc_beta):
    """device memspace (no staging of f) and get/set of the filter state"""
    import torch
    s = H.make_system(3)
    eng, box = _coloured_engine(s, ni_trunc_beta, 7, 1, 5e-4)
    fx = O.Fix(s, O.Beta(path=ni_trunc_beta), O.FDM(2, 2, 2, box, 300.0, 3.5e-6, 1.0, 0.1248), 7, dt=1e-4)
    fx.set_colour(5e-4)
    xis = [np.random.default_rng(i).normal(size=(s["nlocal"], 3)) for i in range(2)]
    recs = traj.run_engine(eng, s, xis, [58.71], 1e-4, device=True, coloured=True)
    refs = traj.run_oracle(fx, s, xis, [58.71])
    for a, b in zip(recs, refs):
        for k in ("f", "f_dis", "f_sto", "T"):
            assert H.error_metrics(a[k], b[k]) < TOL, k
    fd, fs = eng.colour_state()
    eng.set_colour_state(np.ascontiguousarray(2.0 * fd), np.ascontiguousarray(3.0 * fs))
    fd2, fs2 = eng.colour_state()
    assert np.array_equal(fd2, 2.0 * fd) and np.array_equal(fs2, 3.0 * fs)
    assert torch.cuda.is_available()


def test_fix_coloured_b200_matches_committed_golden_vectors():
    g = np.load(os.path.join(GOLDEN, "coloured_case.npz"))
    s = traj.system_from_golden(g)
    args = H.fix_args(int(g["flags"]), os.path.join(GOLDEN, "Ni_trunc.beta"), ["Ni"], model=repr(float(g["tau0"])), grid=(2, 2, 2),
                      group="bit1", style="eph/coloured/exp/b200", extra=["rng", "mars"])
    drv = host.FixDriver(s, args, dt=float(g["dt"]))
    recs = traj.run_fix_driver(drv, s, list(g["xi"]), vec3_probes=dict(f_dis=5, f_sto=6))
    for k in ("f", "array", "T", "w", "rho", "x", "v", "f_dis", "f_sto"):
        got = np.array([r[k] for r in recs])
        assert H.error_metrics(got, g["out_" + k]) < TOL, k
    for k in ("Ee", "Tmean"):
        got = np.array([r[k] for r in recs])
        assert np.all(np.abs(got - g["out_" + k]) <= TOL * np.abs(g["out_" + k])), k
