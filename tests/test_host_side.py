"""Host side of the product (C++ under user-eph_b200/fix): table construction, file grammars,
the C-ABI library's symbols, and the loud failure without a GPU.  No compute calls here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from eph_harness import harness as H
from eph_b200 import host, lib
from oracle import oracle as O

from conftest import REFERENCE, ROOT, gpu_available


def test_product_tables_match_oracle_bit_exact(synth_beta_4, ni_trunc_beta):
    for path in (synth_beta_4, ni_trunc_beta):
        pb, ob = host.BetaTables(path=path), O.Beta(path=path)
        assert (pb.n_elements, pb.n_rho, pb.n_beta) == (ob.n_elements, ob.n_rho, ob.n_beta)
        for a in ("r_cutoff", "r_cutoff_sq", "rho_cutoff", "inv_dr", "inv_dr_sq", "inv_drho"):
            assert getattr(pb, a) == getattr(ob, a), a
        for kind in range(4):
            assert np.array_equal(pb.table(kind), ob.table(kind)), kind
    assert [host.BetaTables(path=synth_beta_4).name(e) for e in range(4)] == ["Ni", "Co", "Cr", "Fe"]


@pytest.mark.parametrize("rel", ["Data/Ni/Ni_PRB2019.beta", "Data/NiCoCrFe/NiCoCrFe_PRB2019.beta", "Examples/Beta/Ni.beta",
                                 "Data/Si/Si_PRB2021_constant.beta"])
def test_product_tables_match_reference_on_shipped_files(ref, rel):
    path = os.path.join(REFERENCE, rel)
    if not os.path.exists(path):
        pytest.skip("reference data tree not present")
    pb, rb = host.BetaTables(path=path), ref.beta_tables(path)
    for kind in range(4):
        assert np.array_equal(pb.table(kind), rb.table(kind)), kind
    assert pb.rho_cutoff == rb.rho_cutoff and pb.inv_dr_sq == rb.inv_dr_sq


def test_spline_builder_matches_oracle():
    rng = np.random.default_rng(21)
    for n in (5, 6, 64, 1001):
        y = rng.normal(size=n)
        assert np.array_equal(host.spline_build(0.37, y), O.spline_build(0.37, y))
    y = np.array([0, 0, 0, 0, 1, 1, 1, 1, 5, 5, 0, 0, 0, 0.0])
    assert np.array_equal(host.spline_build(1.0, y), O.spline_build(1.0, y))


def test_knots_and_file_paths_agree(tmp_path):
    knots = H.synthetic_knots(2, n_rho=301, n_beta=801, drho=0.02)
    p = H.write_beta_file(tmp_path / "k.beta", knots)
    a, b = host.BetaTables(path=p), host.BetaTables(knots=knots)
    for kind in range(4):
        assert np.array_equal(a.table(kind), b.table(kind))


def test_missing_beta_file_is_an_error(tmp_path):
    with pytest.raises(RuntimeError):
        host.BetaTables(path=tmp_path / "nope.beta")


def test_grid_file_grammar_and_writers(tmp_path):
    rng = np.random.default_rng(22)
    nT, dT = 101, 50.0
    par = H.write_parameter_file(tmp_path / "par.data", dT, 3.5e-6 * (1 + np.arange(nT) / 50.0), 0.1 + 0.001 * np.arange(nT))
    nx, ny, nz = 3, 4, 2
    n = nx * ny * nz
    fl = rng.integers(0, 3, n)
    td = rng.integers(0, 2, n)
    path = H.write_grid_file(tmp_path / "T.in", nx, ny, nz, [0, 3, -1, 3, 2, 4], 300 + rng.random(n), rng.random(n), 1 + rng.random(n),
                             3.5e-6 * (1 + rng.random(n)), 0.1 * (1 + rng.random(n)), fl, td, steps=5, parameter_file=str(par))
    g, o = host.GridFile(path), O.FDM(path=path)
    assert (g.nx, g.ny, g.nz, g.steps, g.n_T) == (nx, ny, nz, 5, nT)
    for which in range(5):
        assert np.array_equal(g.field(which), o.field(which))
    assert np.array_equal(g.flags()[0], o.flags()[0]) and np.array_equal(g.flags()[1], o.flags()[1])
    # writers are byte-compatible with the reference grammar (as restated in the oracle)
    g.write_heat_map(g.field(0), tmp_path / "prod_T", 3)
    o.save_temperature(tmp_path / "orc_T", 3)
    assert open(tmp_path / "prod_T_000003").read() == open(tmp_path / "orc_T_000003").read()
    g.write_restart(g.field(0), tmp_path / "prod.restart")
    o.save_state(tmp_path / "orc.restart")
    assert open(tmp_path / "prod.restart").read() == open(tmp_path / "orc.restart").read()
    # and a restart file can be read back as a grid file
    g2 = host.GridFile(tmp_path / "prod.restart")
    assert (g2.nx, g2.ny, g2.nz, g2.steps) == (nx, ny, nz, 5)


def test_reference_example_grid_files_parse(has_reference_tree):
    if not has_reference_tree:
        pytest.skip("reference tree not present")
    cwd = os.getcwd()
    os.chdir(os.path.join(REFERENCE, "Examples/Example_6"))   # T.in names Parameters.data relative to the run directory
    try:
        g, o = host.GridFile("T.in"), O.FDM(path="T.in")
    finally:
        os.chdir(cwd)
    assert (g.nx, g.ny, g.nz) == (9, 9, 9) and g.n_T == 1001
    for which in range(5):
        assert np.array_equal(g.field(which), o.field(which))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "eph_b200.h")).read()
    declared = set(re.findall(r"\b(eph_b200_[A-Za-z_0-9]+)\s*\(", hdr))
    assert len(declared) >= 25
    L = lib.load()
    for name in sorted(declared):
        assert hasattr(L, name), name
    assert declared == set(lib.SYMBOLS), declared ^ set(lib.SYMBOLS)
    assert lib.load().eph_b200_version() == 100


def test_no_cpu_fallback_create_fails_loudly_without_gpu():
    if gpu_available():
        pytest.skip("a GPU is visible")
    with pytest.raises(lib.EphError) as e:
        lib.Engine([0], flags=7)
    assert "no CUDA device" in str(e.value) or "CUDA" in str(e.value)


def test_bad_config_is_rejected_before_touching_the_device():
    L = lib.load()
    h = C.c_void_p()
    tm = (C.c_int * 1)(0)
    cfg = lib.Config(0, 1, tm, 1, 7, 3, 1, 0, 1, None)   # model 3 (PRLCM) is not offered (reference bug, fix_eph.cpp:601)
    assert L.eph_b200_create(C.byref(cfg), C.byref(h)) == -4
    assert b"model" in L.eph_b200_create_error()
    cfg = lib.Config(0, 0, tm, 1, 7, 4, 1, 0, 1, None)
    assert L.eph_b200_create(C.byref(cfg), C.byref(h)) == -1


def test_fix_b200_without_gpu_reports_through_lammps_error(sys500, synth_beta_1):
    """Same constructor checks and messages as the reference (fix_eph.cpp:64, :175, :196); then a loud device error."""
    s = sys500
    with pytest.raises(host.FixError, match="too few arguments"):
        host.FixDriver(s, ["fx", "all", "eph", 1, 7, 4])
    with pytest.raises(host.FixError, match="elements not found"):
        host.FixDriver(s, H.fix_args(7, synth_beta_1, ["Xx"]))
    with pytest.raises(host.FixError, match="model 3 .PRLCM. is not offered"):
        host.FixDriver(s, H.fix_args(7, synth_beta_1, ["Ni"], model=3))
    with pytest.raises(host.FixError, match="unknown model"):
        host.FixDriver(s, H.fix_args(7, synth_beta_1, ["Ni"], model=9))
    with pytest.raises(host.FixError, match="non-positive grid"):
        host.FixDriver(s, H.fix_args(7, synth_beta_1, ["Ni"], grid=(0, 1, 1)))
    for extra, msg in ((["peratom", -1], "peratom must be >= 0"), (["rng", "mt"], "rng must be mars or philox"),
                       (["neigh", "host"], "neigh must be device or lammps"), (["comm", "mpi"], "comm must be device, lammps or nccl"),
                       (["grid", "split"], "grid must be replicated or sharded")):
        with pytest.raises(host.FixError, match=msg):
            host.FixDriver(s, H.fix_args(7, synth_beta_1, ["Ni"], style="eph/b200", extra=extra))
    # decks list more element names than atom types (`Ni.beta Ni Ni`): a keyword is found behind any number of extras
    with pytest.raises(host.FixError, match="peratom must be >= 0"):
        host.FixDriver(s, H.fix_args(7, synth_beta_1, ["Ni", "Ni"], style="eph/b200", extra=["peratom", -1]))
    with pytest.raises(host.FixError, match="keyword without a value"):
        host.FixDriver(s, H.fix_args(7, synth_beta_1, ["Ni", "Ni", "Ni"], style="eph/b200", extra=["rng"]))
    if not gpu_available():
        with pytest.raises(host.FixError, match="no CUDA device|CUDA"):
            host.FixDriver(s, H.fix_args(7, synth_beta_1, ["Ni"]))


def _deck_fix_lines():
    """every `fix ... eph ...` command of the reference's examples, tests and benchmarks (model 4 decks)"""
    out = []
    if not os.path.isdir(REFERENCE):
        return out
    seen = set()
    for top in ("Examples", "Tests", "TB_Bench"):
        for dirpath, _, files in os.walk(os.path.join(REFERENCE, top)):
            for fn in files:
                if not fn.endswith(".lmp"):
                    continue
                for line in open(os.path.join(dirpath, fn), errors="replace"):
                    w = line.split()
                    if len(w) > 18 and w[0] == "fix" and w[3] in ("eph", "eph/gpu") and tuple(w[4:]) not in seen:
                        seen.add(tuple(w[4:]))
                        out.append((os.path.relpath(dirpath, REFERENCE), w))
    return out


@pytest.mark.parametrize("deck", _deck_fix_lines(), ids=lambda d: d[0].replace("/", "_"))
def test_reference_decks_are_accepted_up_to_the_device(deck, sys500):
    """Drop-in check of the command line: each distinct `fix eph` line the reference ships is parsed by FixEPHB200 in
    the deck's own directory (beta file, grid file) and fails -- on a box without a GPU -- only at eph_b200_create."""
    rel, w = deck
    beta = os.path.join(REFERENCE, rel, w[17])
    if not os.path.exists(beta):
        pytest.skip("deck needs %s, which the reference does not ship" % w[17])
    args = ["fx", "all", "eph/b200"] + w[4:]
    cwd = os.getcwd()
    os.chdir(os.path.join(REFERENCE, rel))
    try:
        if gpu_available():
            host.FixDriver(sys500, args).close()
        else:
            with pytest.raises(host.FixError, match="no CUDA device|CUDA"):
                host.FixDriver(sys500, args)
    finally:
        os.chdir(cwd)


def test_harness_neighbor_list_is_a_full_list(sys500):
    s = sys500
    off, ne, x = s["offsets"], s["neigh"], s["x"]
    assert s["nlocal"] == 500 and off[-1] == len(ne)
    nn = np.diff(off)
    assert 120 < nn.mean() < 150          # N_nb ~ 135 at r_c + skin = 7 A (SURVEY.md conventions)
    i = 17
    d = np.linalg.norm(x - x[i], axis=1)
    want = np.nonzero((d < 7.0) & (np.arange(len(x)) != i))[0]
    assert np.array_equal(np.sort(ne[off[i]:off[i + 1]]), want)
    # in-cutoff pairs: 54 for perfect fcc Ni at r_c = 5 A; the 4th shell (4.978 A) straddles the cut-off once displaced
    nc = np.mean([(np.linalg.norm(x[ne[off[k]:off[k + 1]]] - x[k], axis=1) < 5.0).sum() for k in range(50)])
    assert 45 < nc < 55
    # ghosts are images of their owners
    sh = x[s["nlocal"]:] - x[s["ghost_owner"]]
    assert np.allclose(np.abs(sh / s["box"]).round(), np.abs(sh / s["box"]), atol=1e-12)


def test_bench_and_entry_points_parse_without_a_gpu():
    """bench.py's argument surface (the driver's contract flags) and the graft entry module import on a CPU-only box."""
    import subprocess
    import sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--help"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0
    for flag in ("--gpus", "--steps", "--warmup", "--impl"):
        assert flag in out.stdout
    import importlib
    sys.path.insert(0, ROOT)
    g = importlib.import_module("__graft_entry__")
    assert callable(g.build) and callable(g.smoke)


# ---- `fix eph/atomic` boundary (include/eph_b200_atomic.h) ----
def test_atomic_library_exports_every_declared_symbol():
    from eph_b200 import atomic as A
    hdr = open(os.path.join(ROOT, "include", "eph_b200_atomic.h")).read()
    declared = set(re.findall(r"\b(eph_b200_atomic_[A-Za-z_0-9]+)\s*\(", hdr))
    assert len(declared) >= 15
    L = A.load()
    for name in sorted(declared):
        assert hasattr(L, name), name
    assert declared == set(A.SYMBOLS), declared ^ set(A.SYMBOLS)


def test_atomic_no_cpu_fallback_and_argument_errors():
    from eph_b200 import atomic as A
    from eph_harness import harness as H
    if not gpu_available():
        with pytest.raises(lib.EphError) as e:
            A.AtomicEngine([0], [0], 7)
        assert "no CUDA device" in str(e.value) or "CUDA" in str(e.value)
    # the host class parses the reference's command line up to the creation of the device engine
    s = H.make_system(2)
    kappa = os.path.join(ROOT, "tests", "golden", "synth1.kappa")
    beta = os.path.join(ROOT, "tests", "golden", "Ni_trunc.beta")
    good = H.atomic_fix_args(7, beta, kappa, ["Ni"], style="eph/atomic/b200")
    for bad, msg in ((good[:11], "too few arguments"), (good[:-1] + ["Xx"], "elements not found"),
                     (good[:9] + ["/nonexistent.beta"] + good[10:], "cannot open beta file"),
                     (good[:10] + ["/nonexistent.kappa"] + good[11:], "cannot open kappa file"),
                     (good + ["rng", "bad"], "rng must be")):
        with pytest.raises(host.FixError) as e:
            A.fix_driver(s, bad)
        assert msg in str(e.value), (msg, str(e.value))
    if not gpu_available():
        with pytest.raises(host.FixError) as e:
            A.fix_driver(s, good)
        assert "no CUDA device" in str(e.value) or "CUDA" in str(e.value)


def test_atomic_kappa_tables_match_oracle_bit_for_bit():
    from eph_b200 import atomic as A
    from oracle import oracle as O
    path = os.path.join(ROOT, "tests", "golden", "synth1.kappa")
    kp, ko = A.KappaTables(path), O.Kappa(path)
    for a in ("n_elements", "n_pairs", "n_r", "n_T", "r_cutoff", "r_cutoff_sq", "T_max", "inv_dr_sq", "dT"):
        assert getattr(kp, a) == getattr(ko, a), a
    for kind in range(4):
        assert np.array_equal(kp.table(kind), ko.table(kind)), kind
    assert kp.name(0) == "Ni"
    q = np.linspace(0.0, 900.0, 37)
    assert np.array_equal(kp.linear(0, q), O.linear_eval(ko.dT, ko.table(2), q))
    E = O.linear_eval(ko.dT, ko.table(2), q)
    assert np.array_equal(kp.linear(0, E, reverse=True), O.linear_eval(ko.dT, ko.table(2), E, reverse=True))


def test_coloured_fix_command_line_without_a_gpu(sys500, ni_trunc_beta):
    """`fix eph/coloured/exp/b200`: same argument positions as the reference fork, arg[5] = tau0 (fix_eph_coloured_exp.cpp:43)"""
    good = H.fix_args(7, ni_trunc_beta, ["Ni"], model="5e-4", grid=(2, 2, 2), style="eph/coloured/exp/b200")
    with pytest.raises(host.FixError, match="tau0 must be positive"):
        host.FixDriver(sys500, H.fix_args(7, ni_trunc_beta, ["Ni"], model="0", grid=(2, 2, 2), style="eph/coloured/exp/b200"))
    with pytest.raises(host.FixError, match="too few arguments"):
        host.FixDriver(sys500, good[:17])
    if gpu_available():
        host.FixDriver(sys500, good).close()
    else:
        with pytest.raises(host.FixError, match="no CUDA device|CUDA"):
            host.FixDriver(sys500, good)
