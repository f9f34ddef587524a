"""Pins the C restatement (oracle/eph_oracle.c) against the UNMODIFIED reference
compiled into oracle/_ref/libeph_ref.so, bit for bit, and against the golden
files the reference's own tests hold for this path."""
import os

import numpy as np
import pytest

from eph_harness import harness as H
from oracle import oracle as O

import traj
from conftest import REFERENCE

NI = os.path.join(REFERENCE, "Data/Ni/Ni_PRB2019.beta")
HEA = os.path.join(REFERENCE, "Data/NiCoCrFe/NiCoCrFe_PRB2019.beta")


def test_spline_coefficients_bit_exact(ref):
    rng = np.random.default_rng(1)
    for n, dx in ((5, 0.3), (17, 0.01), (1001, 0.005)):
        y = rng.normal(size=n)
        assert np.array_equal(O.spline_build(dx, y), ref.spline_build(dx, y))
    # flat stretches exercise the equal-slope branches of the tangent rule (eph_spline.h:98-106)
    y = np.array([0, 0, 0, 1, 2, 3, 3, 3, 3, 1, 0, 0, 0], dtype=float)
    assert np.array_equal(O.spline_build(0.5, y), ref.spline_build(0.5, y))
    # sin(x) on 1001 points as in the reference's Tests/EPH_Spline/test.cpp:13-51
    xs = np.linspace(0, 10, 1001)
    k = O.spline_build(xs[1], np.sin(xs))
    assert np.array_equal(k, ref.spline_build(xs[1], np.sin(xs)))
    q = np.random.default_rng(2).random(200) * 9.99
    vals = O.spline_eval(k, 1.0 / xs[1], q)
    assert np.array_equal(vals, ref.spline_eval(xs[1], np.sin(xs), q))
    assert np.max(np.abs(vals - np.sin(q))) < 1e-6


def test_linear_table_bit_exact(ref):
    y = np.cumsum(np.random.default_rng(3).random(50))
    q = np.random.default_rng(4).random(100) * 4.8
    assert np.array_equal(O.linear_eval(0.1, y, q), ref.linear_eval(0.1, y, q))
    qy = y[0] + np.random.default_rng(5).random(100) * (y[-2] - y[0])
    assert np.array_equal(O.linear_eval(0.1, y, qy, reverse=True), ref.linear_eval(0.1, y, qy, reverse=True))


@pytest.mark.parametrize("path", [NI, HEA, os.path.join(REFERENCE, "Tests/EPH_Beta/NiFe.beta")])
def test_beta_tables_bit_exact_on_reference_data(ref, path):
    if not os.path.exists(path):
        pytest.skip("reference data tree not present")
    rb, ob = ref.beta_tables(path), O.Beta(path=path)
    assert (rb.n_elements, rb.n_rho, rb.n_beta) == (ob.n_elements, ob.n_rho, ob.n_beta)
    for a in ("r_cutoff", "r_cutoff_sq", "rho_cutoff", "inv_dr", "inv_dr_sq", "inv_drho"):
        assert getattr(rb, a) == getattr(ob, a)
    for kind in range(4):
        assert np.array_equal(rb.table(kind), ob.table(kind)), kind
    if path.endswith("NiFe.beta"):   # expected header of the reference's Tests/EPH_Beta/test.cpp:13-21
        assert rb.n_elements == 2 and rb.r_cutoff == 5.0 and rb.rho_cutoff == 10.0


def test_beta_lookup_bit_exact(ref, synth_beta_4):
    rb, ob = ref.beta_tables(synth_beta_4), O.Beta(path=synth_beta_4)
    rng = np.random.default_rng(6)
    r2 = rng.random(300) * ob.r_cutoff_sq * 0.9999
    rho = np.concatenate([rng.random(300) * ob.rho_cutoff, [ob.rho_cutoff * 1.5]])   # above rho_cutoff -> 0
    for e in range(4):
        assert np.array_equal(ob.eval(1, e, r2), ref.beta_eval(rb, 1, e, r2))
        assert np.array_equal(ob.eval(2, e, rho), ref.beta_eval(rb, 2, e, rho))
        assert np.array_equal(ob.eval(3, e, rho), ref.beta_eval(rb, 3, e, rho))
    assert ob.eval(2, 0, [ob.rho_cutoff * 1.5])[0] == 0.0


def _fdm_pair(ref, nx, ny, nz, box, rng, walls=False, constant=False):
    kw = dict(T_e=300.0, C_e=3.5e-6, rho_e=1.0, kappa_e=0.1248)
    r, o = ref.FDM(nx, ny, nz, box, **kw), O.FDM(nx, ny, nz, box, **kw)
    n = nx * ny * nz
    T = 300 + 100 * rng.random(n)
    kap = 0.1248 * (0.5 + rng.random(n))
    Ce = 3.5e-6 * (0.5 + rng.random(n))
    S = 1e-3 * rng.random(n)
    fl = np.ones(n, dtype=np.int16)
    if walls:
        fl[rng.random(n) < 0.15] = 2
    if constant:
        fl[rng.random(n) < 0.1] = 0
    for which, val in ((0, T), (1, S), (3, Ce), (4, kap)):
        r.set(which, val)
        o.field(which)[:] = val
    r.set_flags(flag=fl)
    o.flags()[0][:] = fl
    return r, o


@pytest.mark.parametrize("shape,walls,constant", [((8, 1, 1), False, False), ((5, 4, 3), True, True), ((1, 1, 1), False, False),
                                                   ((6, 6, 2), True, False)])
def test_fdm_solve_bit_exact(ref, shape, walls, constant):
    rng = np.random.default_rng(11)
    box = [0.0, 17.6, -1.0, 16.6, 2.0, 19.6]
    r, o = _fdm_pair(ref, *shape, box, rng, walls, constant)
    pts = np.stack([rng.uniform(-20, 40, 200), rng.uniform(-20, 40, 200), rng.uniform(-20, 40, 200)], axis=1)
    assert np.array_equal(r.index(pts), np.array([o.index(*p) for p in pts]))   # incl. negative / wrapped coordinates
    for dt in (1e-4, 5e-3):   # the second one needs sub-steps (r > 0.4)
        r.set_dt(dt)
        o.set_dt(dt)
        for _ in range(3):
            E = rng.normal(size=50) * 1e-3
            r.insert_energy(pts[:50], E)
            o.insert_energy(pts[:50], E)
            assert np.array_equal(r.get(5), o.field(5))
            r.solve()
            o.solve()
            assert np.array_equal(r.get(0), o.field(0))
        assert np.array_equal(r.get_T(pts), o.get_T(pts))
        assert r.T_total() == o.T_total()


def test_fdm_files_and_temperature_dependent_cells(ref, tmp_path):
    rng = np.random.default_rng(12)
    nT, dT = 401, 25.0
    Tt = np.arange(nT) * dT
    par = H.write_parameter_file(tmp_path / "par.data", dT, 3.5e-6 * (1 + Tt / 3000.0), 0.1248 * (1 + Tt / 5000.0))
    nx, ny, nz = 4, 3, 3
    n = nx * ny * nz
    tdyn = (rng.random(n) < 0.5).astype(int)
    grid = H.write_grid_file(tmp_path / "T.in", nx, ny, nz, [0, 10, 0, 9, 0, 8], 300 + 2000 * rng.random(n), 0.0, 1.0,
                             3.5e-6, 0.1248, 1, tdyn, steps=3, parameter_file=str(par))
    r, o = ref.FDM(path=grid), O.FDM(path=grid)
    for which in range(5):
        assert np.array_equal(r.get(which), o.field(which))
    assert np.array_equal(r.get_flags()[1], o.flags()[1])
    r.set_dt(2e-4)
    o.set_dt(2e-4)
    pts = rng.uniform(0, 8, (30, 3))
    for _ in range(4):
        E = rng.normal(size=30) * 1e-2
        r.insert_energy(pts, E)
        o.insert_energy(pts, E)
        r.solve()
        o.solve()
        assert np.array_equal(r.get(0), o.field(0))
        assert np.array_equal(r.get(3), o.field(3)) and np.array_equal(r.get(4), o.field(4))
    # writers: heat map and restart, byte for byte
    r.save_temperature(str(tmp_path / "ref_T"), 7)
    o.save_temperature(str(tmp_path / "orc_T"), 7)
    assert open(tmp_path / "ref_T_000007").read() == open(tmp_path / "orc_T_000007").read()
    r.save_state(str(tmp_path / "ref.restart"))
    o.save_state(str(tmp_path / "orc.restart"))
    assert open(tmp_path / "ref.restart").read() == open(tmp_path / "orc.restart").read()


@pytest.mark.parametrize("flags", [1, 3, 7, 2 | 4, 7 | 16, 7 | 32])
def test_fix_hot_path_bit_exact(ref, sys500, synth_beta_1, flags):
    s = sys500
    rng = np.random.default_rng(13)
    xis = [rng.normal(size=(s["nlocal"], 3)) for _ in range(2)]
    drv = ref.fix_driver(s, H.fix_args(flags, synth_beta_1, ["Ni"], grid=(3, 2, 2)), dt=1e-4)
    box = [0, s["box"][0], 0, s["box"][1], 0, s["box"][2]]
    fx = O.Fix(s, O.Beta(path=synth_beta_1), O.FDM(3, 2, 2, box, 300.0, 3.5e-6, 1.0, 0.1248), flags, dt=1e-4)
    a = traj.run_fix_driver(drv, s, xis)
    b = traj.run_oracle(fx, s, xis, [58.71])
    for ra, rb in zip(a, b):
        for k in ("x", "v", "f", "array", "T", "w"):
            assert np.array_equal(ra[k], rb[k]), k
        assert np.array_equal(ra["rho"][: s["nlocal"]], rb["rho"][: s["nlocal"]])
        assert ra["Ee"] == rb["Ee"]


@pytest.mark.parametrize("model", [1, 2])
@pytest.mark.parametrize("flags", [1, 2 | 4, 7])
def test_fix_legacy_models_bit_exact(ref, sys500, synth_beta_1, model, flags):
    """TTM (1, fix_eph.cpp:468-503) and PRB (2, :505-568)."""
    s = sys500
    rng = np.random.default_rng(15)
    xis = [rng.normal(size=(s["nlocal"], 3)) for _ in range(2)]
    drv = ref.fix_driver(s, H.fix_args(flags, synth_beta_1, ["Ni"], model=model, grid=(3, 2, 2)), dt=1e-4)
    box = [0, s["box"][0], 0, s["box"][1], 0, s["box"][2]]
    fx = O.Fix(s, O.Beta(path=synth_beta_1), O.FDM(3, 2, 2, box, 300.0, 3.5e-6, 1.0, 0.1248), flags, model=model, dt=1e-4)
    a = traj.run_fix_driver(drv, s, xis)
    b = traj.run_oracle(fx, s, xis, [58.71])
    for ra, rb in zip(a, b):
        for k in ("x", "v", "f", "array", "T", "w"):
            assert np.array_equal(ra[k], rb[k]), k
        assert ra["Ee"] == rb["Ee"]
    assert np.abs(b[-1]["f"]).max() > 0


@pytest.mark.parametrize("model", [1, 2])
def test_fix_legacy_models_multi_element_bit_exact(ref, synth_beta_4, model):
    s = H.make_system(4, ntypes=3, group_fraction=0.5, pos_seed=5)
    rng = np.random.default_rng(16)
    xis = [rng.normal(size=(s["nlocal"], 3)) for _ in range(2)]
    drv = ref.fix_driver(s, H.fix_args(7, synth_beta_4, ["Fe", "Ni", "Cr"], model=model, grid=(2, 2, 2), group="bit1"),
                         dt=1e-4, mass=[55.85, 58.71, 52.0])
    box = [0, s["box"][0], 0, s["box"][1], 0, s["box"][2]]
    fx = O.Fix(s, O.Beta(path=synth_beta_4), O.FDM(2, 2, 2, box, 300.0, 3.5e-6, 1.0, 0.1248), 7, model=model, groupbit=2,
               type_map=[3, 0, 2], dt=1e-4)
    a = traj.run_fix_driver(drv, s, xis)
    b = traj.run_oracle(fx, s, xis, [55.85, 58.71, 52.0])
    for ra, rb in zip(a, b):
        for k in ("f", "array", "T", "w"):
            assert np.array_equal(ra[k], rb[k]), k


def test_fix_multi_element_group_bit_exact(ref, synth_beta_4):
    s = H.make_system(4, ntypes=3, group_fraction=0.5, pos_seed=5)
    rng = np.random.default_rng(14)
    xis = [rng.normal(size=(s["nlocal"], 3)) for _ in range(2)]
    drv = ref.fix_driver(s, H.fix_args(7, synth_beta_4, ["Fe", "Ni", "Cr"], grid=(2, 2, 2), group="bit1"), dt=1e-4,
                         mass=[55.85, 58.71, 52.0])
    box = [0, s["box"][0], 0, s["box"][1], 0, s["box"][2]]
    ob = O.Beta(path=synth_beta_4)   # element order in the file: Ni Co Cr Fe
    fx = O.Fix(s, ob, O.FDM(2, 2, 2, box, 300.0, 3.5e-6, 1.0, 0.1248), 7, groupbit=2, type_map=[3, 0, 2], dt=1e-4)
    a = traj.run_fix_driver(drv, s, xis)
    b = traj.run_oracle(fx, s, xis, [55.85, 58.71, 52.0])
    for ra, rb in zip(a, b):
        for k in ("f", "array", "T", "w"):
            assert np.array_equal(ra[k], rb[k]), k


@pytest.mark.parametrize("name", ["caseA_example1", "caseB_grid", "caseC_alloy_group"])
def test_oracle_reproduces_committed_golden_vectors(name, ni_trunc_beta):
    """The golden files were generated from the compiled reference (tests/golden/make_golden.py)."""
    from conftest import GOLDEN
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    s = traj.system_from_golden(g)
    box = [0, s["box"][0], 0, s["box"][1], 0, s["box"][2]]
    dt = float(g["dt"])
    if name == "caseA_example1":
        fx = O.Fix(s, O.Beta(path=ni_trunc_beta), O.FDM(1, 1, 1, box, 300.0, 3.5e-6, 1.0, 0.1248), 3, dt=dt)
        mass = [58.71]
    elif name == "caseB_grid":
        cwd = os.getcwd()
        os.chdir(GOLDEN)
        try:
            fdm = O.FDM(path="caseB_grid.in")
        finally:
            os.chdir(cwd)
        fx = O.Fix(s, O.Beta(path=ni_trunc_beta), fdm, 7, dt=dt)
        mass = [58.71]
    else:
        ob = O.Beta(path=os.path.join(GOLDEN, "synth2.beta"))   # elements in file: Ni Co ; fix maps type1->Co, type2->Ni
        fx = O.Fix(s, ob, O.FDM(2, 2, 2, box, 300.0, 3.5e-6, 1.0, 0.1248), 7, groupbit=2, type_map=[1, 0], dt=dt)
        mass = [58.93, 58.71]
    recs = traj.run_oracle(fx, s, list(g["xi"]), mass)
    for k, r in enumerate(recs):
        for key in ("f", "array", "T", "w", "x", "v"):
            assert np.array_equal(r[key], g["out_" + key][k]), (key, k)
        assert r["Ee"] == g["out_Ee"][k]


@pytest.mark.parametrize("test", ["Test1", "Test2"])
def test_reference_fdm_golden_files_loose(test):
    """The reference's own FDM goldens (Tests/EPH_FDM/TestN/Out_Ref/T_out_*): a 1000x1x1 grid heated by a delta
    source for 1 ps, then diffusing (Test1/test.cpp; Test2 adds C_e(x), Test2/test.cpp:101-103).  They were written
    by an older solver with 7 digits, so they pin the restated solver only loosely (SURVEY.md section 4 measured
    ~1e-4 early, 1e-6 later); the 1e-10 pins are the bit-exact tests above."""
    base = os.path.join(REFERENCE, "Tests/EPH_FDM", test, "Out_Ref")
    if not os.path.isdir(base):
        pytest.skip("reference test tree not present")
    n, dt, Q = 1000, 0.001, 10.0
    c_e = 2.0 if test == "Test1" else 1.0
    o = O.FDM(n, 1, 1, [-10.0, 10.0, -1.0, 1.0, -1.0, 1.0], T_e=1.0, C_e=c_e, rho_e=1.0, kappa_e=2.0)
    if test == "Test2":
        o.field(3)[:] = 2.0 * c_e + c_e * np.sin(np.arange(n) * 2.0 * np.pi / n)
    o.set_dt(dt)
    dV = (20.0 / n) * 2.0 * 2.0
    for i in range(0, 8001):
        o.field(1)[:] = 0.0
        if i * dt < 1.0:
            o.field(1)[n // 2] = Q / dV      # cell whose lower corner is x = 0 (Test1/test.cpp:58-63)
        o.solve()
        if i % 1000 == 0 and i // 1000 in (0, 1, 2, 4, 8):
            gold = np.loadtxt(os.path.join(base, "T_out_%06d" % (i // 1000)), skiprows=1)
            assert len(gold) == n
            err = np.max(np.abs(o.field(0) - gold[:, 3]) / np.abs(gold[:, 3]))
            assert err < 2e-3, (i, err)


# ---- `fix eph/coloured/exp` (fix_eph_coloured_exp.cpp): model 4 with an exponential memory kernel on both forces ----
@pytest.fixture(scope="module")
def refc():
    from oracle import reference
    if not reference.coloured_available():
        pytest.skip("oracle/_ref/libeph_coloured_ref.so not built (needs /root/reference at build time)")
    return reference


@pytest.mark.parametrize("flags,group_fraction", [(7, None), (3, None), (1, None), (2, None), (7 + 16, None), (7 + 32, None), (7, 0.6)])
def test_coloured_exp_trajectory_bit_exact(refc, ni_trunc_beta, flags, group_fraction):
    s = H.make_system(3, group_fraction=group_fraction)
    group, gb = ("bit1", 2) if group_fraction else ("all", 1)
    box = [0, s["box"][0], 0, s["box"][1], 0, s["box"][2]]
    tau0 = 5e-4
    drv = refc.coloured_fix_driver(s, H.fix_args(flags, ni_trunc_beta, ["Ni"], model=repr(tau0), grid=(2, 2, 2), group=group,
                                                 style="eph/coloured/exp"))
    fx = O.Fix(s, O.Beta(path=ni_trunc_beta), O.FDM(2, 2, 2, box, 300.0, 3.5e-6, 1.0, 0.1248), flags, groupbit=gb, dt=1e-4)
    fx.set_colour(tau0)
    xis = [np.random.default_rng(i).normal(size=(s["nlocal"], 3)) if flags & 2 else None for i in range(4)]
    recs = traj.run_fix_driver(drv, s, xis, vec3_probes=dict(f_dis=5, f_sto=6))
    refs = traj.run_oracle(fx, s, xis, [58.71])
    for step, (a, b) in enumerate(zip(recs, refs)):
        for k in ("x", "v", "f", "array", "T", "Ee", "Tmean", "w", "rho", "f_dis", "f_sto"):
            assert np.array_equal(np.asarray(a[k]), np.asarray(b[k])), (step, k)
    # the filter has memory: the first step carries zeta times the unfiltered force
    if flags & 1:
        assert np.abs(recs[0]["f_dis"]).max() > 0
    # a time-step change refreshes zeta (fix_eph_coloured_exp.cpp:686)
    drv.set_dt(2e-4)
    fx.set_dt(2e-4)
    a = traj.run_fix_driver(drv, s, xis[:1], vec3_probes=dict(f_dis=5, f_sto=6))[0]
    b = traj.run_oracle(fx, s, xis[:1], [58.71])[0]
    for k in ("f", "f_dis", "f_sto"):
        assert np.array_equal(np.asarray(a[k]), np.asarray(b[k])), k


def test_coloured_oracle_matches_committed_golden_vectors(ni_trunc_beta):
    from conftest import GOLDEN
    g = np.load(os.path.join(GOLDEN, "coloured_case.npz"))
    s = traj.system_from_golden(g)
    box = [0, s["box"][0], 0, s["box"][1], 0, s["box"][2]]
    fx = O.Fix(s, O.Beta(path=ni_trunc_beta), O.FDM(2, 2, 2, box, 300.0, 3.5e-6, 1.0, 0.1248), int(g["flags"]),
               groupbit=int(g["groupbit"]), dt=float(g["dt"]))
    fx.set_colour(float(g["tau0"]))
    recs = traj.run_oracle(fx, s, list(g["xi"]), [58.71])
    for k in ("f", "array", "T", "Ee", "Tmean", "w", "rho", "x", "v", "f_dis", "f_sto"):
        assert np.array_equal(np.array([r[k] for r in recs]), g["out_" + k]), k


def test_reference_coloured_survives_atom_reordering(refc, ni_trunc_beta):
    """the reference fork migrates f_sto_i / f_dis_i with the atoms (fix_eph_coloured_exp.cpp:793-825): the stand-in's
    re-ordering is transparent to it, bit for bit"""
    s = H.make_system(3)
    xi = [np.random.default_rng(60 + k).normal(size=(s["natoms"], 3)) for k in range(4)]
    args = H.fix_args(7, ni_trunc_beta, ["Ni"], model="5e-4", grid=(2, 2, 2), style="eph/coloured/exp")
    traj.assert_reordering_is_transparent(lambda system: refc.coloured_fix_driver(system, args), s, xi, permute_after=2)
