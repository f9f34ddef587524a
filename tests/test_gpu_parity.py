"""Parity of the CUDA path against the oracle on the same seeded inputs, through the C ABI
(eph_b200.lib.Engine) and through the FixEPHB200 host class.  Bar (BASELINE.json north_star):
rho, beta, forces and T_e grids within 1e-10 relative; force errors are scaled by the largest
reference magnitude because per-atom friction is a cancelling sum."""
import os

import numpy as np
import pytest

from eph_harness import harness as H
from eph_b200 import host, lib
from oracle import oracle as O

import traj
from conftest import GOLDEN

pytestmark = pytest.mark.gpu
TOL = 1e-10


def make_engine(beta_path, flags, grid, box, type_map=(0,), groupbit=1, dt=1e-4, grid_file=None, seed=12345, model=4):
    eng = lib.Engine(list(type_map), flags=flags, model=model, groupbit=groupbit, seed=seed)
    eng.set_tables_from(host.BetaTables(path=beta_path))
    if grid_file is not None:
        host.GridFile(grid_file).apply(eng)
    else:
        eng.set_grid(grid[0], grid[1], grid[2], box, 300.0, 1.0, 3.5e-6, 0.1248)
    eng.set_dt(dt)
    return eng


def attach(eng, s, device=False):
    if device:
        import torch
        d = torch.device("cuda", 0)
        t = lambda a, dt: torch.as_tensor(np.ascontiguousarray(a), dtype=dt, device=d)
        eng.set_atoms(s["nlocal"], s["nghost"], t(s["type"], torch.int32), t(s["mask"], torch.int32), t(s["tag"], torch.int64),
                      t(s["ghost_owner"], torch.int32))
        eng.set_neighbors(t(s["offsets"], torch.int64), t(s["neigh"], torch.int32))
    else:
        eng.set_atoms(s["nlocal"], s["nghost"], np.ascontiguousarray(s["type"], dtype=np.int32),
                      np.ascontiguousarray(s["mask"], dtype=np.int32), np.ascontiguousarray(s["tag"], dtype=np.int64),
                      np.ascontiguousarray(s["ghost_owner"], dtype=np.int32))
        eng.set_neighbors(np.ascontiguousarray(s["offsets"], dtype=np.int64), np.ascontiguousarray(s["neigh"], dtype=np.int32))


def compare(recs, refs, nl, keys=("f", "array", "T", "w", "x", "v"), dt=1e-4):
    e_scale = 0.0
    for k, (a, b) in enumerate(zip(recs, refs)):
        for key in keys:
            err = H.error_metrics(a[key], b[key])
            assert err < TOL, (key, k, err)
        assert H.error_metrics(a["rho"][:nl], b["rho"][:nl]) < TOL
        # Ee accumulates -(f_EPH + f_RNG) . v dt over atoms and steps: a cancelling sum (the random force heats, the friction
        # cools), so like the forces it is measured against the size of its terms, not against what is left of them
        e_scale += traj.energy_scale(b, dt)
        assert abs(a["Ee"] - b["Ee"]) <= TOL * max(abs(b["Ee"]), e_scale, 1e-300), (a["Ee"], b["Ee"])
        assert abs(a["Tmean"] - b["Tmean"]) <= TOL * abs(b["Tmean"])


def box6(s):
    return [0, s["box"][0], 0, s["box"][1], 0, s["box"][2]]


@pytest.mark.parametrize("flags", [1, 3, 7, 2 | 4, 7 | 16, 7 | 32])
@pytest.mark.parametrize("device", [False, True])
def test_engine_matches_oracle_flags(sys500, synth_beta_1, flags, device):
    s = sys500
    rng = np.random.default_rng(31)
    xis = [rng.normal(size=(s["nlocal"], 3)) for _ in range(3)]
    fx = O.Fix(s, O.Beta(path=synth_beta_1), O.FDM(3, 2, 2, box6(s), 300.0, 3.5e-6, 1.0, 0.1248), flags, dt=1e-4)
    refs = traj.run_oracle(fx, s, xis, [58.71])
    eng = make_engine(synth_beta_1, flags, (3, 2, 2), box6(s))
    attach(eng, s, device)
    recs = traj.run_engine(eng, s, xis, [58.71], 1e-4, device=device)
    compare(recs, refs, s["nlocal"])
    for a, b in zip(recs, refs):
        assert H.error_metrics(a["f_eph"], b["f_eph"]) < TOL and H.error_metrics(a["f_rng"], b["f_rng"]) < TOL
    assert eng.launch_count() > 0


@pytest.mark.parametrize("lanes", ["1", "2", "8", "16"])
def test_engine_lane_widths(sys500, synth_beta_1, lanes, monkeypatch):
    """every sub-warp width of the sweeps (default 8) and both table paths give the same answer"""
    import subprocess, sys, textwrap
    code = textwrap.dedent("""
        import sys, os, numpy as np
        sys.path.insert(0, %r); sys.path.insert(0, %r); sys.path.insert(0, %r)
        import test_gpu_parity as T, traj
        from eph_harness import harness as H
        from oracle import oracle as O
        s = H.make_system(5)
        xis = [np.random.default_rng(3).normal(size=(s["nlocal"], 3))]
        fx = O.Fix(s, O.Beta(path=%r), O.FDM(2, 2, 2, T.box6(s), 300.0, 3.5e-6, 1.0, 0.1248), 7, dt=1e-4)
        refs = traj.run_oracle(fx, s, xis, [58.71])
        eng = T.make_engine(%r, 7, (2, 2, 2), T.box6(s)); T.attach(eng, s)
        T.compare(traj.run_engine(eng, s, xis, [58.71], 1e-4), refs, s["nlocal"])
        print("ok")
    """) % (os.path.join(os.path.dirname(GOLDEN), "..", "user-eph_b200"), os.path.join(os.path.dirname(GOLDEN), ".."),
            os.path.dirname(GOLDEN), synth_beta_1, synth_beta_1)
    env = dict(os.environ, EPH_B200_LANES=lanes, EPH_B200_TABLE="0" if lanes == "16" else "1")
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "ok" in out.stdout, out.stdout + out.stderr


@pytest.mark.parametrize("model", [1, 2])
@pytest.mark.parametrize("flags", [1, 2 | 4, 7])
def test_engine_legacy_models_match_oracle(sys500, synth_beta_1, model, flags):
    """TTM (fix_eph.cpp:468-503) and PRB (:505-568) on the device against the restatement (itself bit-exact against the
    compiled reference, test_oracle_vs_reference.py::test_fix_legacy_models_bit_exact)."""
    s = sys500
    rng = np.random.default_rng(33)
    xis = [rng.normal(size=(s["nlocal"], 3)) for _ in range(3)]
    fx = O.Fix(s, O.Beta(path=synth_beta_1), O.FDM(3, 2, 2, box6(s), 300.0, 3.5e-6, 1.0, 0.1248), flags, model=model, dt=1e-4)
    refs = traj.run_oracle(fx, s, xis, [58.71])
    eng = make_engine(synth_beta_1, flags, (3, 2, 2), box6(s), model=model)
    attach(eng, s)
    recs = traj.run_engine(eng, s, xis, [58.71], 1e-4)
    compare(recs, refs, s["nlocal"])
    for a, b in zip(recs, refs):
        assert H.error_metrics(a["f_eph"], b["f_eph"]) < TOL and H.error_metrics(a["f_rng"], b["f_rng"]) < TOL
    assert np.abs(refs[-1]["f"]).max() > 0


@pytest.mark.parametrize("model", [1, 2])
def test_engine_legacy_models_multi_element_and_group(synth_beta_4, model):
    s = H.make_system(4, ntypes=3, group_fraction=0.5, pos_seed=5)
    rng = np.random.default_rng(34)
    xis = [rng.normal(size=(s["nlocal"], 3)) for _ in range(2)]
    mass = [55.85, 58.71, 52.0]
    fx = O.Fix(s, O.Beta(path=synth_beta_4), O.FDM(2, 2, 2, box6(s), 300.0, 3.5e-6, 1.0, 0.1248), 7, model=model, groupbit=2,
               type_map=[3, 0, 2], dt=1e-4)
    refs = traj.run_oracle(fx, s, xis, mass)
    eng = make_engine(synth_beta_4, 7, (2, 2, 2), box6(s), type_map=[3, 0, 2], groupbit=2, model=model)
    attach(eng, s)
    compare(traj.run_engine(eng, s, xis, mass, 1e-4), refs, s["nlocal"])


def test_engine_multi_element_and_group(synth_beta_4):
    s = H.make_system(4, ntypes=3, group_fraction=0.5, pos_seed=5)
    rng = np.random.default_rng(32)
    xis = [rng.normal(size=(s["nlocal"], 3)) for _ in range(2)]
    mass = [55.85, 58.71, 52.0]
    fx = O.Fix(s, O.Beta(path=synth_beta_4), O.FDM(2, 2, 2, box6(s), 300.0, 3.5e-6, 1.0, 0.1248), 7, groupbit=2,
               type_map=[3, 0, 2], dt=1e-4)
    refs = traj.run_oracle(fx, s, xis, mass)
    eng = make_engine(synth_beta_4, 7, (2, 2, 2), box6(s), type_map=[3, 0, 2], groupbit=2)
    attach(eng, s)
    compare(traj.run_engine(eng, s, xis, mass, 1e-4), refs, s["nlocal"])


def test_rho_above_cutoff_gives_zero_coupling(tmp_path):
    """the `beta set to zero` branch (eph_beta.h:174-180; reference deck Tests/EPH_Beta_zero)"""
    knots = H.synthetic_knots(1, n_beta=11, drho=0.01)      # rho_cutoff = 0.1, inside the lattice's site-density spread
    p = str(H.write_beta_file(tmp_path / "low.beta", knots))
    s = H.make_system(4, sigma=0.08)
    xi = [np.random.default_rng(33).normal(size=(s["nlocal"], 3))]
    ob = O.Beta(path=p)
    fx = O.Fix(s, ob, O.FDM(1, 1, 1, box6(s), 300.0, 3.5e-6, 1.0, 0.1248), 7, dt=1e-4)
    refs = traj.run_oracle(fx, s, xi, [58.71])
    above = refs[0]["rho"][: s["nlocal"]] > ob.rho_cutoff
    assert above.any() and not above.all()
    eng = make_engine(p, 7, (1, 1, 1), box6(s))
    attach(eng, s)
    compare(traj.run_engine(eng, s, xi, [58.71], 1e-4), refs, s["nlocal"])
    assert eng.status_word() & 1


def test_builtin_gaussian_stream_matches_its_definition(sys500, synth_beta_1):
    """xi from the built-in counter-based stream == the stream's CPU definition fed through the injection port"""
    s = sys500
    eng = make_engine(synth_beta_1, 7, (2, 2, 2), box6(s), seed=777)
    attach(eng, s)
    recs = traj.run_engine(eng, s, [None, None], [58.71], 1e-4)
    tags = s["tag"][: s["nlocal"]]
    xis = [O.xi_stream(777, step, tags) for step in (1, 2)]
    for r, xi in zip(recs, xis):
        assert np.max(np.abs(r["xi"] - xi)) < 1e-13
    fx = O.Fix(s, O.Beta(path=synth_beta_1), O.FDM(2, 2, 2, box6(s), 300.0, 3.5e-6, 1.0, 0.1248), 7, dt=1e-4)
    compare(recs, traj.run_oracle(fx, s, xis, [58.71]), s["nlocal"])
    allxi = np.concatenate([r["xi"].ravel() for r in recs])
    assert abs(allxi.mean()) < 0.1 and abs(allxi.std() - 1.0) < 0.05


@pytest.mark.parametrize("name", ["caseA_example1", "caseB_grid", "caseC_alloy_group"])
def test_engine_matches_committed_golden_vectors(name, ni_trunc_beta):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    s = traj.system_from_golden(g)
    dt = float(g["dt"])
    if name == "caseA_example1":
        eng, mass = make_engine(ni_trunc_beta, 3, (1, 1, 1), box6(s), dt=dt), [58.71]
    elif name == "caseB_grid":
        eng, mass = make_engine(ni_trunc_beta, 7, None, None, dt=dt, grid_file=os.path.join(GOLDEN, "caseB_grid.in")), [58.71]
    else:
        eng = make_engine(os.path.join(GOLDEN, "synth2.beta"), 7, (2, 2, 2), box6(s), type_map=[1, 0], groupbit=2, dt=dt)
        mass = [58.93, 58.71]
    attach(eng, s)
    recs = traj.run_engine(eng, s, list(g["xi"]), mass, dt)
    refs = [dict((k, g["out_" + k][i]) for k in ("f", "array", "T", "w", "x", "v", "rho", "Ee", "Tmean")) for i in range(len(recs))]
    compare(recs, refs, s["nlocal"])


@pytest.mark.parametrize("comm", ["device", "lammps"])
@pytest.mark.parametrize("name", ["caseA_example1", "caseB_grid", "caseC_alloy_group"])
def test_fix_b200_matches_committed_golden_vectors(name, comm):
    """FixEPHB200 in the LAMMPS stand-in, same command line as the reference fix, rng mars = injected stream;
    ghost values either through the engine's owner map or through Comm::forward_comm(Fix*) like the reference"""
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    s = traj.system_from_golden(g)
    s["natoms"] = s["nlocal"]
    dt = float(g["dt"])
    cwd = os.getcwd()
    os.chdir(GOLDEN)
    try:
        if name == "caseA_example1":
            drv = host.FixDriver(s, H.fix_args(3, "Ni_trunc.beta", ["Ni"], grid=(1, 1, 1), style="eph/b200", extra=["rng", "mars", "comm", comm]), dt=dt)
        elif name == "caseB_grid":
            drv = host.FixDriver(s, H.fix_args(7, "Ni_trunc.beta", ["Ni"], T_infile="caseB_grid.in", style="eph/b200", extra=["rng", "mars", "comm", comm]), dt=dt)
        else:
            drv = host.FixDriver(s, H.fix_args(7, "synth2.beta", ["Co", "Ni"], grid=(2, 2, 2), group="bit1", style="eph/b200",
                                               extra=["rng", "mars", "comm", comm]), dt=dt, mass=[58.93, 58.71])
    finally:
        os.chdir(cwd)
    recs = traj.run_fix_driver(drv, s, list(g["xi"]))
    refs = [dict((k, g["out_" + k][i]) for k in ("f", "array", "T", "w", "x", "v", "rho", "Ee", "Tmean")) for i in range(len(recs))]
    compare(recs, refs, s["nlocal"])
    fl = drv.fix_flags()
    assert fl["size_peratom_cols"] == 8 and fl["comm_forward"] == 3 and fl["ghost_velocity"] == 1 and fl["size_vector"] == 2
    assert drv.neigh_cutoff() == 5.0
    if comm == "lammps":      # XI, RHO and (with friction) WI broadcasts per step, as in the reference
        assert drv.n_forward() >= 3 * len(recs)


def _fdm_case(shape, rng, walls, constant, tdyn_tables=None):
    nx, ny, nz = shape
    n = nx * ny * nz
    fl = np.ones(n, dtype=np.int16)
    if walls:
        fl[rng.random(n) < 0.15] = 2
    if constant:
        fl[rng.random(n) < 0.1] = 0
    return dict(T=300 + 100 * rng.random(n), kap=0.1248 * (0.5 + rng.random(n)), Ce=3.5e-6 * (0.5 + rng.random(n)),
                S=1e-3 * rng.random(n), rho=1.0 + 0.2 * rng.random(n), fl=fl)


@pytest.mark.parametrize("shape,walls,constant", [((8, 1, 1), False, False), ((5, 4, 3), True, True), ((1, 1, 1), False, False),
                                                   ((33, 9, 5), True, False), ((64, 64, 64), False, False),
                                                   ((32, 6, 5), True, True), ((48, 16, 12), True, False), ((16, 4, 4), False, False)])
def test_grid_solve_matches_oracle(synth_beta_1, shape, walls, constant):
    """EPH_FDM::solve alone: energy deposited by a handful of atoms, several solves, with and without sub-stepping"""
    rng = np.random.default_rng(41)
    box = [0.0, 17.6, -1.0, 16.6, 2.0, 19.6]
    c = _fdm_case(shape, rng, walls, constant)
    o = O.FDM(*shape, box, 300.0, 3.5e-6, 1.0, 0.1248)
    for which, key in ((0, "T"), (1, "S"), (2, "rho"), (3, "Ce"), (4, "kap")):
        o.field(which)[:] = c[key]
    o.flags()[0][:] = c["fl"]
    eng = lib.Engine([0], flags=7)
    eng.set_tables_from(host.BetaTables(path=synth_beta_1))
    eng.set_grid(*shape, box, c["T"], c["rho"], c["Ce"], c["kap"], S_e=c["S"], flag=c["fl"])
    for dt in ((1e-6, 1e-5) if shape[0] == 64 else (1e-4, 5e-3)):   # the second one needs sub-steps (r > 0.4)
        o.set_dt(dt)
        eng.set_dt(dt)
        for _ in range(3):
            src = 1e-2 * rng.normal(size=o.ntotal)
            o.field(5)[:] = src
            eng.put_grid(5, src)
            o.solve()
            _solve_only(eng)
            assert H.error_metrics(eng.get_grid(0), o.field(0)) < TOL
            assert np.all(eng.get_grid(5) == 0.0)
        assert abs(eng.mean_T() - o.T_total()) < TOL * o.T_total()
    assert eng.last_substeps() > 1


@pytest.mark.parametrize("shape", [(32, 16, 8), (64, 4, 4), (16, 4, 5), (40, 9, 7)])
def test_grid_uniform_fast_path_matches_oracle(synth_beta_1, shape):
    """`NX NY NZ NULL` grids (all cells dynamic, identical parameters) take the constant-coefficient TMA kernel"""
    rng = np.random.default_rng(43)
    box = [0.0, 35.2, 0.0, 17.6, -3.0, 14.6]
    o = O.FDM(*shape, box, 300.0, 3.5e-6, 1.0, 0.1248)
    eng = lib.Engine([0], flags=7)
    eng.set_tables_from(host.BetaTables(path=synth_beta_1))
    eng.set_grid(*shape, box, 300.0, 1.0, 3.5e-6, 0.1248)
    T0 = 300 + 200 * rng.random(o.ntotal)
    o.field(0)[:] = T0
    eng.put_grid(0, T0)                      # T_e is state, not a parameter: the fast path stays on
    for dt in (1e-4, 2e-3):
        o.set_dt(dt)
        eng.set_dt(dt)
        for _ in range(3):
            src = 1e-2 * rng.normal(size=o.ntotal)
            o.field(5)[:] = src
            eng.put_grid(5, src)
            o.solve()
            _solve_only(eng)
            assert H.error_metrics(eng.get_grid(0), o.field(0)) < TOL
            assert np.all(eng.get_grid(5) == 0.0)
    assert eng.last_substeps() > 1


def _solve_only(eng):
    """run end_of_step with no atoms contributing: a one-atom system outside the fix group"""
    if not getattr(eng, "_dummy", False):
        x = np.array([[1.0, 1.0, 3.0]]); z = np.zeros((1, 3))
        eng.set_atoms(1, 0, np.array([1], dtype=np.int32), np.array([0], dtype=np.int32), np.array([1], dtype=np.int64))
        eng.set_neighbors(np.array([0, 0], dtype=np.int64), np.array([0], dtype=np.int32))
        eng._xz = (x, z)
        eng._dummy = True
    x, z = eng._xz
    eng.post_force(x, z, z.copy(), None, 0)
    eng.end_of_step(x, z)


@pytest.mark.parametrize("shape", [(6, 5, 4), (32, 5, 4)])   # generic kernel / TMA-tiled kernel
def test_grid_temperature_dependent_cells(synth_beta_1, tmp_path, shape):
    rng = np.random.default_rng(42)
    nT, dT = 401, 25.0
    Tt = np.arange(nT) * dT
    par = H.write_parameter_file(tmp_path / "par.data", dT, 3.5e-6 * (1 + Tt / 3000.0), 0.1248 * (1 + Tt / 5000.0))
    nx, ny, nz = shape
    n = nx * ny * nz
    grid = H.write_grid_file(tmp_path / "T.in", nx, ny, nz, [0, 10, 0, 9, 0, 8], 300 + 2000 * rng.random(n), 0.0, 1.0, 3.5e-6, 0.1248, 1,
                             (rng.random(n) < 0.5).astype(int), steps=3, parameter_file=str(par))
    o = O.FDM(path=grid)
    eng = lib.Engine([0], flags=7)
    eng.set_tables_from(host.BetaTables(path=synth_beta_1))
    host.GridFile(grid).apply(eng)
    o.set_dt(2e-4)
    eng.set_dt(2e-4)
    for _ in range(4):
        src = 1e-2 * rng.normal(size=n)
        o.field(5)[:] = src
        eng.put_grid(5, src)
        o.solve()
        _solve_only(eng)
        assert H.error_metrics(eng.get_grid(0), o.field(0)) < TOL
        assert H.error_metrics(eng.get_grid(3), o.field(3)) < TOL and H.error_metrics(eng.get_grid(4), o.field(4)) < TOL


@pytest.mark.parametrize("skin", [None, 2.0])
def test_inner_list_invalidation_and_rebuild(synth_beta_1, skin):
    """The two-level Verlet list must never change results: atoms are displaced by more than half the inner skin
    between steps (neighbour list NOT rebuilt, as LAMMPS would not for moves below skin/2), which trips the
    device-side check, falls back to LAMMPS' list and (with a known skin) rebuilds the inner list."""
    s = H.make_system(5)
    nl = s["nlocal"]
    rng = np.random.default_rng(61)
    eng = make_engine(synth_beta_1, 7, (2, 2, 2), box6(s))
    if skin is not None:
        eng.set_skin(skin, 0.4)
    attach(eng, s)
    fdm = O.FDM(2, 2, 2, box6(s), 300.0, 3.5e-6, 1.0, 0.1248)
    fx = O.Fix(s, O.Beta(path=synth_beta_1), fdm, 7, dt=1e-4)
    sync = traj.GhostSync(s)
    x, v = s["x"].copy(), s["v"].copy()
    # displacement schedule: small, small, big (> 0.2 A for some atoms), small, big again, ... total stays below skin/2 = 1 A
    for step, amp in enumerate([0.0, 0.01, 0.25, 0.01, 0.0, 0.3, 0.02, 0.02], start=1):
        x[:nl] += amp * rng.uniform(-1, 1, (nl, 3)) / np.sqrt(3)
        sync(x, v)
        xi = rng.normal(size=(nl, 3))
        fx.x[:] = x; fx.v[:] = v; fx.f[:] = 0.0
        fx.post_force(xi)
        fx.end_of_step()
        f = np.zeros((nl, 3))
        eng.post_force(x, v, f, xi, step)
        eng.end_of_step(x, v)
        assert H.error_metrics(f, fx.f[:nl]) < TOL, step
        assert H.error_metrics(eng.probe(0)[:nl], np.array(fx.ptr(0))[:nl]) < TOL, step
        assert H.error_metrics(eng.get_grid(0), fdm.field(0)) < TOL, step
    st = eng.list_stats()
    assert st["fallback_steps"] >= 1
    assert st["inner_builds"] >= (2 if skin else 1)


@pytest.mark.parametrize("world,overlap", [(2, False), (4, False), (2, True), (4, True)])
def test_bricks_with_ghost_exchange_match_whole_box_oracle(synth_beta_1, world, overlap):
    """The multi-rank path on one GPU: `world` engines, one per spatial brick, driven through the two halves of
    post_force / end_of_step with the ghost payload and the grid source term moved between them exactly as
    torch.distributed would (eph_harness.parallel); the result must equal the single-rank oracle on the whole box."""
    import torch
    from eph_harness import parallel as P
    dev = torch.device("cuda", 0)
    # one stream for torch and all engines, so the copies that stand in for NCCL are ordered with the kernels
    tstream = torch.cuda.Stream(device=dev)
    gstream = torch.cuda.Stream(device=dev)
    # overlap: the exchange runs on a communication stream behind the boundary tiles of the density pass
    # (eph_b200_set_comm_stream / eph_b200_set_boundary_atoms) while the main stream sweeps the interior tiles
    cstream = torch.cuda.Stream(device=dev) if overlap else tstream
    torch.cuda.set_stream(tstream)
    n, seed, dt = 6, 4711, 1e-4
    whole = H.make_system(n)
    grid = P.brick_grid(world)
    systems = [H.make_system(n, brick=(r, grid)) for r in range(world)]
    plans = P.ExchangePlan.build_all(systems)
    gshape = (3, 2, 2)
    box = box6(whole)
    engines, state = [], []
    t = lambda a, ty: torch.as_tensor(np.ascontiguousarray(a), dtype=ty, device=dev)
    for r, (s, plan) in enumerate(zip(systems, plans)):
        eng = lib.Engine([0], flags=7, seed=seed, rank=r, nranks=world, stream=tstream.cuda_stream)
        eng.set_tables_from(host.BetaTables(path=synth_beta_1))
        eng.set_grid(*gshape, box, 300.0, 1.0, 3.5e-6, 0.1248)
        eng.set_dt(dt)
        eng.set_atoms(s["nlocal"], s["nghost"], t(s["type"], torch.int32), t(s["mask"], torch.int32), t(s["tag"], torch.int64),
                      t(plan.self_owner, torch.int32))
        eng.set_neighbors(t(s["offsets"], torch.int64), t(s["neigh"], torch.int32))
        src = torch.zeros(int(np.prod(gshape)), dtype=torch.float64, device=dev)
        eng.bind_grid_source(src)
        eng.set_grid_stream(gstream.cuda_stream)     # source all-reduce + solve on a second stream
        if overlap:
            eng.set_comm_stream(cstream.cuda_stream)
            eng.set_boundary_atoms(t(plan.flat_send_index(), torch.int32))
        engines.append(eng)
        state.append(dict(x=t(s["x"], torch.float64), v=t(s["v"], torch.float64),
                          f=torch.zeros((s["nlocal"], 3), dtype=torch.float64, device=dev), src=src,
                          send=t(plan.flat_send_index(), torch.int32), recv=t(plan.flat_recv_index(), torch.int32)))
    # oracle on the whole box with the same counter-based Gaussians
    fx = O.Fix(whole, O.Beta(path=synth_beta_1), O.FDM(*gshape, box, 300.0, 3.5e-6, 1.0, 0.1248), 7, dt=dt)
    for step in (1, 2):
        xi = O.xi_stream(seed, step, whole["tag"][: whole["nlocal"]])
        fx.f[:] = 0.0
        fx.post_force(xi)
        fx.end_of_step()
        for eng, st in zip(engines, state):
            st["f"].zero_()
            eng.post_force_begin(st["x"], st["v"], None, step)
        # the exchange: all_to_all of {rho, W} rows, done here by plain copies between the engines' buffers
        with torch.cuda.stream(cstream):
            bufs = []
            for eng, st in zip(engines, state):
                b = torch.empty((max(st["send"].numel(), 1), 4), dtype=torch.float64, device=dev)
                if st["send"].numel():
                    eng.pack_ghost_payload(st["send"], b)
                bufs.append(b)
            for r, (eng, st, plan) in enumerate(zip(engines, state, plans)):
                parts = []
                for q in range(world):   # what rank q sends to r sits in q's buffer after what q sends to ranks < r
                    off = sum(plans[q].send_counts[:r])
                    parts.append(bufs[q][off: off + plans[q].send_counts[r]])
                rb = torch.cat(parts) if parts else torch.empty((0, 4), dtype=torch.float64, device=dev)
                assert rb.shape[0] == st["recv"].numel()
                if st["recv"].numel():
                    eng.unpack_ghost_payload(st["recv"], rb.contiguous())
        for eng, st in zip(engines, state):
            eng.post_force_end(st["f"])
            eng.end_of_step_begin(st["x"], st["v"])
        with torch.cuda.stream(gstream):    # all-reduce of the source term, on the grid stream like dist.all_reduce would be
            total = torch.zeros_like(state[0]["src"])
            for st in state:
                total += st["src"]
            for st in state:
                st["src"].copy_(total)
        E = 0.0
        for eng, st in zip(engines, state):
            E += eng.end_of_step_end(True)
        # compare per atom by tag
        ref_f, ref_rho = fx.f[: whole["nlocal"]], np.array(fx.ptr(0))[: whole["nlocal"]]
        order = np.argsort(whole["tag"][: whole["nlocal"]])
        for eng, st, s in zip(engines, state, systems):
            idx = order[np.searchsorted(whole["tag"][: whole["nlocal"]][order], s["tag"][: s["nlocal"]])]
            assert H.error_metrics(st["f"].cpu().numpy(), ref_f[idx], floor=np.abs(ref_f).max()) < TOL
            assert H.error_metrics(eng.probe(0)[: s["nlocal"]], ref_rho[idx]) < TOL
            assert H.error_metrics(eng.get_grid(0), fx.fdm.field(0)) < TOL
        assert abs(E - (fx.Ee() if step == 1 else fx.Ee() - E_prev)) < 1e-9 * abs(E)
        E_prev = fx.Ee()
    torch.cuda.synchronize()
    torch.cuda.set_stream(torch.cuda.default_stream(dev))


def test_device_built_neighbor_list(synth_beta_1):
    """eph_b200_build_neighbors: same pair set as the host list (LAMMPS' REQ_FULL semantics) and same results"""
    s = H.make_system(7, sigma=0.08)
    nl = s["nlocal"]
    eng = make_engine(synth_beta_1, 7, (2, 2, 2), box6(s))
    eng.set_atoms(nl, s["nghost"], np.ascontiguousarray(s["type"], dtype=np.int32), np.ascontiguousarray(s["mask"], dtype=np.int32),
                  np.ascontiguousarray(s["tag"], dtype=np.int64), np.ascontiguousarray(s["ghost_owner"], dtype=np.int32))
    eng.build_neighbors(s["x"], 7.0)
    off, ne = eng.get_neighbors()
    assert np.array_equal(off, s["offsets"])          # same row lengths ...
    for i in range(0, nl, 37):                         # ... and the same members
        assert np.array_equal(np.sort(ne[off[i]:off[i + 1]]), np.sort(s["neigh"][s["offsets"][i]:s["offsets"][i + 1]]))
    xi = [np.random.default_rng(71).normal(size=(nl, 3)) for _ in range(2)]
    fx = O.Fix(s, O.Beta(path=synth_beta_1), O.FDM(2, 2, 2, box6(s), 300.0, 3.5e-6, 1.0, 0.1248), 7, dt=1e-4)
    compare(traj.run_engine(eng, s, xi, [58.71], 1e-4), traj.run_oracle(fx, s, xi, [58.71]), nl)
    # device pointers, and the fix keyword
    import torch
    eng.build_neighbors(torch.as_tensor(s["x"], device="cuda"), 7.0)
    off2, ne2 = eng.get_neighbors()
    assert np.array_equal(off2, off) and np.array_equal(ne2, ne)
    g = np.load(os.path.join(GOLDEN, "caseA_example1.npz"))
    sg = traj.system_from_golden(g)
    sg["natoms"] = sg["nlocal"]
    cwd = os.getcwd()
    os.chdir(GOLDEN)
    try:
        drv = host.FixDriver(sg, H.fix_args(3, "Ni_trunc.beta", ["Ni"], grid=(1, 1, 1), style="eph/b200",
                                            extra=["rng", "mars", "neigh", "device"]), dt=float(g["dt"]))
    finally:
        os.chdir(cwd)
    recs = traj.run_fix_driver(drv, sg, list(g["xi"]))
    refs = [dict((k, g["out_" + k][i]) for k in ("f", "array", "T", "w", "x", "v", "rho", "Ee", "Tmean")) for i in range(len(recs))]
    compare(recs, refs, sg["nlocal"])
    assert drv.neigh_cutoff() == 0.0                   # no list was requested from LAMMPS


def test_empty_and_ragged_inputs(synth_beta_1):
    eng = make_engine(synth_beta_1, 7, (2, 2, 2), [0, 10, 0, 10, 0, 10])
    z = np.zeros((0, 3))
    eng.set_atoms(0, 0, np.zeros(0, dtype=np.int32), np.zeros(0, dtype=np.int32), np.zeros(0, dtype=np.int64))
    eng.set_neighbors(np.zeros(1, dtype=np.int64), np.zeros(1, dtype=np.int32))
    eng.post_force(z, z, z.copy(), None, 1)            # no atoms: nothing to do, no error
    # ragged: atoms with empty rows, isolated atoms (rho = 0 -> skipped everywhere, fix_eph.cpp:709)
    x = np.array([[1.0, 1, 1], [2.5, 1, 1], [9.0, 9, 9], [1.0, 3.2, 1]])
    v = np.random.default_rng(1).normal(size=(4, 3))
    off = np.array([0, 2, 4, 4, 6], dtype=np.int64)
    ne = np.array([1, 3, 0, 3, 0, 1], dtype=np.int32)
    s = dict(nlocal=4, nghost=0, x=x, v=v, f=np.zeros_like(x), type=np.ones(4, dtype=np.int32), mask=np.ones(4, dtype=np.int32),
             tag=np.arange(1, 5), ghost_owner=np.zeros(0, dtype=np.int32), offsets=off, neigh=ne, box=np.array([10.0, 10, 10]), ntypes=1)
    xi = [np.random.default_rng(2).normal(size=(4, 3))]
    fx = O.Fix(s, O.Beta(path=synth_beta_1), O.FDM(2, 2, 2, [0, 10, 0, 10, 0, 10], 300.0, 3.5e-6, 1.0, 0.1248), 7, dt=1e-4)
    refs = traj.run_oracle(fx, s, xi, [58.71])
    attach(eng, s)
    recs = traj.run_engine(eng, s, xi, [58.71], 1e-4)
    compare(recs, refs, 4)
    assert np.all(recs[0]["array"][2] == 0.0)          # the isolated atom: rho = 0, no forces
    # call-order errors are reported, not ignored
    e2 = lib.Engine([0], flags=7)
    with pytest.raises(lib.EphError, match="set_tables"):
        e2.post_force(x, v, np.zeros_like(x), None, 0)


def test_size_independent_properties_at_scale(synth_beta_1):
    """n = 24 (55 296 atoms, 7.5 M list entries): properties that need no oracle run --
    momentum conservation of the pair forces, linearity in v and xi, energy bookkeeping."""
    s = H.make_system(24)
    nl = s["nlocal"]
    eng = make_engine(synth_beta_1, 7 | 16 | 32, (8, 8, 8), box6(s))   # forces computed but not applied
    attach(eng, s)
    rng = np.random.default_rng(51)
    xi = rng.normal(size=(nl, 3))
    x, v = s["x"], s["v"]
    f = np.zeros((nl, 3))
    eng.post_force(x, v, f, xi, 0)    # the step that builds the inner list runs on the fp64 records; the ones below on one kind
    eng.post_force(x, v, f, xi, 1)
    assert np.all(f == 0.0)
    fe, fr, rho = eng.probe(3), eng.probe(4), eng.probe(0)
    assert np.all(rho[:nl] > 0) and np.allclose(rho[nl:], rho[s["ghost_owner"]], rtol=0, atol=0)
    # every pair term is antisymmetric: the total friction and random force vanish
    assert np.max(np.abs(fe.sum(axis=0))) < 1e-9 * np.abs(fe).max() * np.sqrt(nl)
    assert np.max(np.abs(fr.sum(axis=0))) < 1e-9 * np.abs(fr).max() * np.sqrt(nl)
    # friction is linear in v, the random force linear in xi
    eng.post_force(x, 2.0 * v, f, -3.0 * xi, 1)
    # (a factor 2 commutes with the block-floating-point records; a factor -3 re-rounds 40-bit mantissas: <= 2^-39 per vector)
    assert H.error_metrics(eng.probe(3), 2.0 * fe) < 1e-12 and H.error_metrics(eng.probe(4), -3.0 * fr) < 2e-11
    # friction dissipates: -f_EPH . v >= 0 summed over atoms; energy bookkeeping matches the deposited source
    eng.post_force(x, v, f, xi, 1)
    T0 = eng.get_grid(0)
    E = eng.end_of_step(x, v)
    fe, fr = eng.probe(3), eng.probe(4)
    E_ref = -np.sum((fe + fr) * v[:nl]) * 1e-4
    assert abs(E - E_ref) < 1e-10 * max(abs(E_ref), np.abs(fe * v[:nl]).sum() * 1e-4)
    assert -np.sum(fe * v[:nl]) > 0
    # grid energy: sum(C_e rho_e dT) dV equals the energy handed over (periodic grid, uniform parameters)
    dV = np.prod(s["box"] / 8)
    dE_grid = np.sum(eng.get_grid(0) - T0) * 3.5e-6 * 1.0 * dV
    assert abs(dE_grid - E) < 1e-8 * abs(E)
    arr = eng.peratom()
    rho = eng.probe(0)
    assert np.array_equal(arr[:, 0], rho[:nl]) and np.array_equal(arr[:, 2:5], fe) and np.array_equal(arr[:, 5:8], fr)


@pytest.mark.parametrize("case", ["Ni_cascade", "NiCoCrFe_partial_group"])
def test_all_atoms_against_compiled_reference_at_131k_atoms(case, ref):
    """The unmodified reference fix (oracle/_ref, about half a second per step on the host) and FixEPHB200 on the device
    on a 131 072-atom box: the shipped parametrisations (Ni_PRB2019: 50 001 beta knots; NiCoCrFe_PRB2019: four elements),
    the 10 keV primary knock-on atom, a partial fix group, six Verlet steps with one re-neighbouring in the middle (ghosts
    and list rebuilt from the current positions).  ALL atoms are compared: rho_i, w_i, the per-atom output (rho, beta,
    f_EPH, f_RNG), f, x, v, and the whole T_e grid, at 1e-10."""
    from eph_b200 import host
    s = H.make_system(32, ntypes=4 if case != "Ni_cascade" else 1, group_fraction=None if case == "Ni_cascade" else 0.7)
    s["v"][0] = 1813.0 * np.array([0.835115, 0.543981, 0.081652])
    s["v"][s["nlocal"]:][s["ghost_owner"] == 0] = s["v"][0]
    elems, beta = (["Ni"], H.shipped_beta("Ni_PRB2019")) if case == "Ni_cascade" else (["Ni", "Co", "Cr", "Fe"], H.shipped_beta("NiCoCrFe_PRB2019"))
    group = "all" if case == "Ni_cascade" else "bit1"
    dt = 1e-4 if case != "Ni_cascade" else 5.5e-7     # the cascade's adaptive step while the PKA is fast
    xis = [np.random.default_rng(300 + k).normal(size=(s["nlocal"], 3)) for k in range(6)]
    ref_args = H.fix_args(7, beta, elems, grid=(16, 16, 16), group=group)
    # the alloy case lets the engine build the full list itself (`neigh device`: the fill pass also writes the inner list)
    our_args = H.fix_args(7, beta, elems, grid=(16, 16, 16), group=group, style="eph/b200",
                          extra=["rng", "mars"] + ([] if case == "Ni_cascade" else ["neigh", "device"]))
    dts = [dt] * len(xis)
    a = traj.run_with_reneighbouring(lambda sy: ref.fix_driver(sy, ref_args, dt=dt), s, xis, {3: 7.0}, dts=dts, probes=True)
    b = traj.run_with_reneighbouring(lambda sy: host.FixDriver(sy, our_args, dt=dt), s, xis, {3: 7.0}, dts=dts, probes=True)
    traj.assert_same_trajectory(a, b, TOL, dts=dts)
    for step, (ra, rb) in enumerate(zip(a, b), start=1):
        for key in ("rho", "w", "grid"):
            assert H.error_metrics(rb[key], ra[key]) < TOL, (step, key)
        assert np.abs(ra["array"][:, 2:8]).max() > 0 and np.all(ra["rho"][(np.asarray(s["mask"][: s["nlocal"]]) & (2 if group == "bit1" else 1)) != 0] > 0)


def test_full_size_box_properties_and_sampled_parity(ni_trunc_beta):
    """BASELINE's full size (config C3: n = 100, 4 000 000 atoms, 5.4e8 list entries, the bench's .beta file).  The C
    oracle needs minutes per step there, so parity is checked two ways: size-independent properties over all atoms
    (momentum conservation, energy bookkeeping against the grid) and an independent numpy evaluation of rho_i, f_EPH_i
    and f_RNG_i for a sample of atoms -- straight from the definitions (fix_eph.cpp:450-461, :713-739, :758-785,
    :802-833) with the oracle's spline tables, including the neighbours' rho_j and w_j."""
    s = H.make_system(100)
    nl, x, v = s["nlocal"], s["x"], s["v"]
    off, ne, owner = s["offsets"], s["neigh"], s["ghost_owner"]
    dt, Te = 1e-4, 300.0
    eng = make_engine(ni_trunc_beta, 7 | 16 | 32, (64, 64, 64), box6(s), dt=dt)
    attach(eng, s)
    rng = np.random.default_rng(61)
    xi = rng.normal(size=(nl, 3))
    f = np.zeros((nl, 3))
    eng.post_force(x, v, f, xi, 1)
    T0 = eng.get_grid(0)
    E = eng.end_of_step(x, v)
    fe, fr, rho = eng.probe(3), eng.probe(4), eng.probe(0)
    assert np.all(rho[:nl] > 0)
    # properties over all 4 M atoms
    assert np.max(np.abs(fe.sum(axis=0))) < 1e-9 * np.abs(fe).max() * np.sqrt(nl)
    assert np.max(np.abs(fr.sum(axis=0))) < 1e-9 * np.abs(fr).max() * np.sqrt(nl)
    E_ref = -np.sum((fe + fr) * v[:nl]) * dt
    assert abs(E - E_ref) < 1e-10 * max(abs(E_ref), np.abs(fe * v[:nl]).sum() * dt)
    dV = np.prod(s["box"] / 64)
    assert abs(np.sum(eng.get_grid(0) - T0) * 3.5e-6 * 1.0 * dV - E) < 1e-8 * abs(E)

    # sampled parity against the definitions
    ob = O.Beta(path=ni_trunc_beta)
    t_rho, t_alpha = ob.table(1)[0], ob.table(2)[0]
    rc2 = ob.r_cutoff_sq

    def spline(tab, inv_dx, xs):
        k = (xs * inv_dx).astype(np.int64)
        a, b, c, d = tab[k, 0], tab[k, 1], tab[k, 2], tab[k, 3]
        return a + xs * (b + xs * (c + xs * d))

    def home(j):                      # a ghost carries its owner's rho, w and xi
        return j if j < nl else int(owner[j - nl])

    cache = {}

    def site(i):                      # rho_i, s_i = alpha/rho, W_i of a local atom, from its own list
        if i not in cache:
            jj = ne[off[i]:off[i + 1]] & 0x1FFFFFFF
            e = x[jj] - x[i]
            r2 = np.einsum("ij,ij->i", e, e)
            m = r2 < rc2
            e, r2, jj = e[m], r2[m], jj[m]
            rj = spline(t_rho, ob.inv_dr_sq, r2)
            r_i = rj.sum()
            g = rj / r2
            W = ((g * np.einsum("ij,ij->i", e, v[i] - v[jj]))[:, None] * e).sum(axis=0)
            al = 0.0 if r_i > ob.rho_cutoff else float(spline(t_alpha, ob.inv_drho, np.array([r_i]))[0])
            cache[i] = (r_i, al / r_i, W, e, g, jj)
        return cache[i]

    eta = np.sqrt(2.0 * H.KB / dt)
    sample = rng.choice(nl, 16, replace=False)
    for i in sample:
        r_i, s_i, W_i, e, g, jj = site(int(i))
        u_i, z_i = s_i * s_i * W_i, s_i * xi[i]
        fe_i, fr_i = np.zeros(3), np.zeros(3)
        for k, j in enumerate(jj):
            r_j, s_j, W_j = site(home(int(j)))[:3]
            u_j, z_j = s_j * s_j * W_j, s_j * xi[home(int(j))]
            fe_i -= g[k] * (e[k] @ u_i - e[k] @ u_j) * e[k]
            fr_i += g[k] * (e[k] @ z_i - e[k] @ z_j) * e[k]
        fr_i *= eta * np.sqrt(Te)
        assert abs(rho[i] - r_i) < TOL * r_i
        assert np.max(np.abs(fe[i] - fe_i)) < TOL * np.abs(fe).max(), (i, fe[i], fe_i)
        assert np.max(np.abs(fr[i] - fr_i)) < TOL * np.abs(fr).max(), (i, fr[i], fr_i)


def test_fluctuation_dissipation_statistics(synth_beta_1):
    """<f_RNG f_RNG^T> over noise realisations equals 2 k_B T_e B / dt with f_EPH = -B v
    (the random force is built from the same W as the friction: PRL 120, 185501)."""
    s = H.make_system(6)
    nl = s["nlocal"]
    dt, Te = 1e-4, 300.0
    eng = make_engine(synth_beta_1, 7 | 16 | 32, (1, 1, 1), box6(s), dt=dt, seed=4242)
    attach(eng, s)
    x = s["x"]
    f = np.zeros((nl, 3))
    # B column for (atom a, direction d): friction response to a unit velocity of one atom and its images
    a = 17
    B = np.zeros((3, 3))
    for d in range(3):
        v = np.zeros_like(x)
        v[a, d] = 1.0
        v[nl:][s["ghost_owner"] == a, d] = 1.0
        eng.post_force(x, v, f, None, 0)
        B[:, d] = -eng.probe(3)[a]
    # sample the random force on atom a over many steps of the built-in stream
    nsamp = 3000
    acc = np.zeros((3, 3))
    v0 = np.zeros_like(x)
    for step in range(nsamp):
        eng.post_force(x, v0, f, None, step + 1)
        fr = eng.probe(4)[a]
        acc += np.outer(fr, fr)
    cov = acc / nsamp
    want = 2.0 * H.KB * Te * B / dt
    assert np.allclose(B, B.T, rtol=1e-9, atol=1e-12 * np.abs(B).max())
    assert np.max(np.abs(cov - want)) < 0.12 * np.max(np.abs(want))
