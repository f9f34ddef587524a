"""Parity cases of the `fix eph/atomic` device path against the oracle, shared by the `-m gpu` tests (the CUDA build on
a B200) and tests/test_atomic_emulated.py (the same source compiled for the host, tests/emul).  Every function takes
`make_engine(type_map_beta, type_map_kappa, flags, **kw)` returning an eph_b200.atomic.AtomicEngine, and where the host
class is involved `make_fix(system, args)` returning a FixDriver of FixEPHAtomicB200."""
import os

import numpy as np

from eph_harness import harness as H
from eph_b200 import host
from oracle import oracle as O

import traj
from conftest import GOLDEN

KAPPA = os.path.join(GOLDEN, "synth1.kappa")
BETA = os.path.join(GOLDEN, "Ni_trunc.beta")
TOL = 1e-10   # north star: relative, forces and energies scaled by the largest reference magnitude
KEYS = ("f", "array", "rho", "w", "f_eph", "f_rng", "rho_a", "E", "dE", "T", "x", "v")


def setup_engine(eng, s, beta_path, kappa_tables, dt=1e-4, T_init=300.0):
    eng.set_tables_from(host.BetaTables(path=beta_path), kappa_tables)
    eng.set_dt(dt)
    eng.set_atoms(s["nlocal"], s["nghost"], np.ascontiguousarray(s["type"], dtype=np.int32),
                  np.ascontiguousarray(s["mask"], dtype=np.int32), np.ascontiguousarray(s["tag"], dtype=np.int64),
                  np.ascontiguousarray(s["ghost_owner"], dtype=np.int32))
    eng.set_neighbors(np.ascontiguousarray(s["offsets"], dtype=np.int64), np.ascontiguousarray(s["neigh"], dtype=np.int32))
    eng.init_energy(T_init)
    return eng


def compare(recs, refs, in_group=None):
    worst = 0.0
    for step, (a, b) in enumerate(zip(recs, refs)):
        for k in KEYS:
            got, ref = np.asarray(a[k]), np.asarray(b[k])
            if k == "T" and in_group is not None:
                got, ref = got[in_group], ref[in_group]
            assert np.all(np.isfinite(got)), (step, k)
            err = H.error_metrics(got, ref)
            assert err < TOL, (step, k, err)
            worst = max(worst, err)
        for k in ("Ee", "Te"):
            assert abs(a[k] - b[k]) <= TOL * abs(b[k]), (step, k, a[k], b[k])
    return worst


def trajectory_case(make_engine, kappa_tables, flags, loops, group_fraction=None, n=3, steps=3, ntypes=1, beta=BETA,
                    names=("Ni",), kappa=KAPPA, tk=None):
    s = H.make_system(n, group_fraction=group_fraction, ntypes=ntypes)
    gb = 2 if group_fraction else 1
    tm = list(range(ntypes)) if len(names) > 1 else [0] * ntypes
    tk = list(tk) if tk is not None else [0] * ntypes    # element of the .kappa file per atom type
    eng = setup_engine(make_engine(tm, tk, flags, groupbit=gb, inner_loops=loops), s, beta, kappa_tables)
    fx = O.AtomicFix(s, O.Beta(path=beta), O.Kappa(kappa), flags, groupbit=gb, inner_loops=loops, type_map_beta=tm,
                     type_map_kappa=tk)
    # the constructor's sums (fix_eph_atomic.cpp:223-253)
    Ee0, Te0 = eng.summary()
    assert abs(Ee0 - fx.Ee()) <= TOL * abs(fx.Ee()) and abs(Te0 - fx.Te()) <= TOL * abs(fx.Te())
    xis = [np.random.default_rng(10 + i).normal(size=(s["nlocal"], 3)) if flags & 2 else None for i in range(steps)]
    mass = [58.71] * ntypes
    recs = traj.run_atomic_engine(eng, s, xis, mass, 1e-4, groupbit=gb, noint=bool(flags & 8))
    refs = traj.run_atomic_oracle(fx, s, xis, mass)
    in_group = (s["mask"][: s["nlocal"]] & gb) != 0
    worst = compare(recs, refs, in_group)
    assert eng.launch_count() > 0
    return worst


def gradient_case(make_engine, kappa_tables):
    """heat diffusion alone (flags 4) from an energy gradient, 4 inner loops"""
    s = H.make_system(3)
    eng = setup_engine(make_engine([0], [0], 4, inner_loops=4), s, BETA, kappa_tables)
    fx = O.AtomicFix(s, O.Beta(path=BETA), O.Kappa(KAPPA), 4, inner_loops=4)
    E0 = np.array(fx.ptr(6)[: s["nlocal"]]) * (1.0 + 0.8 * np.sin(2 * np.pi * s["x"][: s["nlocal"], 0] / s["box"][0]))
    eng.set_energy(np.ascontiguousarray(E0))
    fx.set_energy(E0)
    recs = traj.run_atomic_engine(eng, s, [None] * 4, [58.71], 1e-4)
    refs = traj.run_atomic_oracle(fx, s, [None] * 4, [58.71])
    compare(recs, refs)
    assert np.std(recs[-1]["T"]) < np.std(recs[0]["T"])


def golden_engine_case(make_engine, kappa_tables, name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    s = traj.system_from_golden(g)
    gb = int(g["groupbit"])
    eng = setup_engine(make_engine([0], [0], int(g["flags"]), groupbit=gb, inner_loops=int(g["inner_loops"])), s, BETA,
                       kappa_tables, dt=float(g["dt"]))
    if "E0" in g.files:
        eng.set_energy(np.ascontiguousarray(g["E0"]))
    recs = traj.run_atomic_engine(eng, s, list(g["xi"]), [58.71], float(g["dt"]), groupbit=gb)
    for k in ("f", "array", "rho", "w", "f_eph", "f_rng", "rho_a", "E", "dE", "x", "v"):
        got = np.array([r[k] for r in recs])
        assert H.error_metrics(got, g["out_" + k]) < TOL, k
    for k in ("Ee", "Te"):
        got = np.array([r[k] for r in recs])
        assert np.all(np.abs(got - g["out_" + k]) <= TOL * np.abs(g["out_" + k])), k


def golden_fix_case(make_fix, name, comm="device"):
    """the committed golden vectors through FixEPHAtomicB200 (constructor syntax, hooks, outputs); comm = "lammps" sends
    the ghost values through Comm::forward_comm(Fix*) and the fix's pack/unpack callbacks like the reference"""
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    s = traj.system_from_golden(g)
    group = "bit1" if int(g["groupbit"]) == 2 else "all"
    args = H.atomic_fix_args(int(g["flags"]), BETA, KAPPA, ["Ni"], inner_loops=int(g["inner_loops"]), group=group,
                             style="eph/atomic/b200") + ["rng", "mars", "comm", comm]
    drv = make_fix(s, args)
    n0 = drv.n_forward()
    fl = drv.fix_flags()
    assert fl["size_peratom_cols"] == 12 and fl["comm_forward"] == 3 and fl["ghost_velocity"] == 1 and drv.neigh_cutoff() == 5.0
    if "E0" in g.files:
        drv.set_energy(np.ascontiguousarray(g["E0"]))
    recs = traj.run_atomic_fix_driver(drv, s, list(g["xi"]))
    for k in ("f", "array", "rho", "w", "f_eph", "f_rng", "rho_a", "E", "dE", "x", "v"):
        got = np.array([r[k] for r in recs])
        assert H.error_metrics(got, g["out_" + k]) < TOL, k
    for k in ("Ee", "Te"):
        got = np.array([r[k] for r in recs])
        assert np.all(np.abs(got - g["out_" + k]) <= TOL * np.abs(g["out_" + k])), k
    # forward comms: the reference's per step (EI, XI, RHO, WI and one EI per heat loop, fix_eph_atomic.cpp:803-825,
    # :549-550, :720-721) with comm lammps; with comm device only the owner map, once per registration (the stand-in's
    # list always counts as fresh, neighbor->ago == 0, so that is once per step here)
    loops = max(int(g["inner_loops"]), 1)
    want = len(recs) * (4 + loops) if comm == "lammps" else len(recs)
    assert drv.n_forward() - n0 == want, (drv.n_forward() - n0, want)


def philox_case(make_engine, kappa_tables):
    """built-in Gaussian stream: xi of a step is the pinned Philox definition keyed on (seed, tag, step), and the run
    equals one with that stream injected"""
    s = H.make_system(3)
    a = setup_engine(make_engine([0], [0], 7, seed=777, inner_loops=1), s, BETA, kappa_tables)
    b = setup_engine(make_engine([0], [0], 7, seed=777, inner_loops=1), s, BETA, kappa_tables)
    tags = np.ascontiguousarray(s["tag"][: s["nlocal"]], dtype=np.int64)
    xis = [O.xi_stream(777, step, tags) for step in (1, 2)]
    ra = traj.run_atomic_engine(a, s, [None, None], [58.71], 1e-4)   # None with RANDOM set -> built-in stream
    rb = traj.run_atomic_engine(b, s, xis, [58.71], 1e-4)
    for x, y, xi in zip(ra, rb, xis):
        assert H.error_metrics(x["xi"], xi) < 1e-13
        for k in ("f", "E", "dE", "f_rng"):
            assert H.error_metrics(x[k], y[k]) < 1e-12, k


def properties_case(make_engine, kappa_tables, n, steps=2, loops=2):
    """Size-independent properties of the path (no oracle needed, any box size): the pair forces are antisymmetric, so
    both forces sum to zero over the box; the energy ledger returns exactly the work the two forces do on the atoms,
    sum_i dE_i = -dt sum_i (f_EPH_i + f_RNG_i).v_i; heat diffusion only moves energy between atoms; rho, rho_a > 0."""
    s = H.make_system(n)
    nl = s["nlocal"]
    eng = setup_engine(make_engine([0], [0], 7, inner_loops=loops), s, BETA, kappa_tables)
    sync = traj.GhostSync(s)
    x = np.ascontiguousarray(s["x"]).copy()
    v = np.ascontiguousarray(s["v"]).copy()
    dt = 1e-4
    E_prev = eng.get_energy()
    for step in range(1, steps + 1):
        f = np.zeros((nl, 3))
        xi = np.random.default_rng(100 + step).normal(size=(nl, 3))
        eng.post_force(x, v, f, xi, step)
        f_eph, f_rng, dE, rho, rho_a = eng.probe(3), eng.probe(4), eng.probe(7), eng.probe(0), eng.probe(5)
        assert np.all(rho[:nl] > 0) and np.all(rho_a[:nl] > 0)
        assert np.array_equal(rho[nl:], rho[sync.owner]) and np.array_equal(rho_a[nl:], rho_a[sync.owner])
        assert H.error_metrics(f, f_eph + f_rng) < 1e-14
        for force in (f_eph, f_rng):
            assert np.max(np.abs(force.sum(axis=0))) < 1e-10 * np.abs(force).sum()
        work = -dt * np.einsum("ij,ij->", f_eph + f_rng, v[:nl])
        assert abs(dE.sum() - work) < 1e-10 * np.abs(dE).sum(), (dE.sum(), work)
        Ee, Te = eng.end_of_step()
        E = eng.get_energy()
        assert np.all(E >= 0)
        assert abs(Ee - E.sum()) < 1e-12 * E.sum()
        assert abs(E.sum() - (E_prev.sum() + dE.sum())) < 1e-10 * E.sum()      # diffusion conserves, nothing was clamped
        assert abs(Te - eng.probe(8).mean()) < 1e-12 * Te
        E_prev = E
        v[:nl] += 0.5 * dt * H.FTM2V / 58.71 * f       # keep the state moving between the steps
        x[:nl] += dt * v[:nl]
        sync(x, v)


def reordering_case(make_fix, comm="device"):
    """LAMMPS re-orders the local atoms (spatial sort): FixEPHAtomicB200 carries E_a_i along through copy_arrays and
    re-registers it on the device, so the trajectory continues as if nothing had happened"""
    s = H.make_system(3)
    xi = [np.random.default_rng(40 + k).normal(size=(s["natoms"], 3)) for k in range(4)]
    args = H.atomic_fix_args(7, BETA, KAPPA, ["Ni"], inner_loops=2, style="eph/atomic/b200") + ["rng", "mars", "comm", comm]
    traj.assert_reordering_is_transparent(lambda system: make_fix(system, args), s, xi, permute_after=2, tol=TOL)


def ragged_case(make_engine, kappa_tables):
    """empty and ragged inputs: no atoms at all; then an isolated cluster without ghosts in which one atom has no
    neighbour (rho = 0: skipped everywhere, fix_eph_atomic.cpp:515) and one lies outside the fix group"""
    eng = make_engine([0], [0], 7, inner_loops=2)
    eng.set_tables_from(host.BetaTables(path=BETA), kappa_tables)
    eng.set_dt(1e-4)
    z = np.zeros((0, 3))
    eng.set_atoms(0, 0, np.zeros(0, dtype=np.int32), np.zeros(0, dtype=np.int32), np.zeros(0, dtype=np.int64), None)
    eng.set_neighbors(np.zeros(1, dtype=np.int64), np.zeros(1, dtype=np.int32))
    eng.post_force(z, z, z.copy(), None, 1)       # nothing to do, no error
    rng = np.random.default_rng(5)
    nl = 20
    x = np.ascontiguousarray(rng.uniform(0, 9, (nl, 3)))
    x[7] = [40.0, 40.0, 40.0]
    v = rng.normal(size=(nl, 3))
    off, ne = H.neighbor_list(x, nl, 7.0)
    mask = np.ones(nl, dtype=np.int32)
    mask[3] = 0
    s = dict(nlocal=nl, nghost=0, x=x, v=v, f=np.zeros_like(x), type=np.ones(nl, dtype=np.int32), mask=mask, tag=np.arange(1, nl + 1),
             ghost_owner=np.zeros(0, dtype=np.int32), offsets=off, neigh=ne, box=np.array([50.0, 50.0, 50.0]), ntypes=1)
    eng.set_atoms(nl, 0, s["type"], mask, np.ascontiguousarray(s["tag"], dtype=np.int64), None)
    eng.set_neighbors(np.ascontiguousarray(off, dtype=np.int64), np.ascontiguousarray(ne, dtype=np.int32))
    eng.init_energy(300.0)
    fx = O.AtomicFix(s, O.Beta(path=BETA), O.Kappa(KAPPA), 7, inner_loops=2)
    xis = [rng.normal(size=(nl, 3)) for _ in range(3)]
    recs = traj.run_atomic_engine(eng, s, xis, [58.71], 1e-4)
    compare(recs, traj.run_atomic_oracle(fx, s, xis, [58.71]), in_group=mask != 0)
    assert recs[-1]["rho"][7] == 0.0 and not recs[-1]["f"][7].any() and not recs[-1]["array"][3].any()
