"""`fix eph/atomic`: pins the C restatement (oracle/eph_oracle_atomic.c) bit for bit against the
UNMODIFIED reference fix_eph_atomic.cpp / eph_kappa.h compiled into oracle/_ref/libeph_atomic_ref.so, and against the
committed golden vectors generated from it (tests/golden/make_golden_atomic.py)."""
import os

import numpy as np
import pytest

from eph_harness import harness as H
from oracle import oracle as O

import traj
from conftest import GOLDEN, REFERENCE

KAPPA = os.path.join(GOLDEN, "synth1.kappa")
BETA = os.path.join(GOLDEN, "Ni_trunc.beta")
KEYS = ("f", "array", "Ee", "Te", "rho", "w", "f_eph", "f_rng", "rho_a", "E", "dE", "x", "v")


@pytest.fixture(scope="module")
def refa():
    from oracle import reference
    if not reference.atomic_available():
        pytest.skip("oracle/_ref/libeph_atomic_ref.so not built (needs /root/reference at build time)")
    return reference


@pytest.mark.parametrize("path", [KAPPA, os.path.join(REFERENCE, "Tests/EPH_kappa/Cu.kappa"),
                                  os.path.join(REFERENCE, "Tests/EPH_atomic_run/heat_diffusion/Ni.kappa")])
def test_kappa_tables_bit_exact(refa, path):
    if not os.path.exists(path):
        pytest.skip("reference data tree not present")
    kr, ko = refa.Kappa(path), O.Kappa(path)
    for a in ("n_elements", "n_pairs", "n_r", "n_T", "r_cutoff", "r_cutoff_sq", "T_max", "inv_dr_sq", "dT"):
        assert getattr(kr, a) == getattr(ko, a), a
    for kind in range(4):
        assert np.array_equal(kr.table(kind), ko.table(kind)), kind
    if path.endswith("Cu.kappa"):   # the one expectation of the reference's Tests/EPH_kappa/test.cpp:17
        assert kr.n_elements == 1


CASES = [(7, 0, None), (7, 3, None), (1, 0, None), (2, 0, None), (3, 0, None), (5, 2, None), (6, 1, None),
         (7 + 16, 2, None), (7 + 32, 2, None), (7 + 8, 1, None), (7, 2, 0.7)]


@pytest.mark.parametrize("flags,loops,group_fraction", CASES)
def test_atomic_trajectory_bit_exact(refa, flags, loops, group_fraction):
    s = H.make_system(3, group_fraction=group_fraction)
    group, gb = ("bit1", 2) if group_fraction else ("all", 1)
    drv = refa.atomic_fix_driver(s, H.atomic_fix_args(flags, BETA, KAPPA, ["Ni"], inner_loops=loops, group=group))
    fx = O.AtomicFix(s, O.Beta(path=BETA), O.Kappa(KAPPA), flags, groupbit=gb, inner_loops=loops)
    assert drv.fix_flags()["size_peratom_cols"] == 12 and drv.neigh_cutoff() == 5.0
    xis = [np.random.default_rng(i).normal(size=(s["nlocal"], 3)) if flags & 2 else None for i in range(3)]
    a = traj.run_atomic_fix_driver(drv, s, xis)
    b = traj.run_atomic_oracle(fx, s, xis, [58.71])
    in_group = (s["mask"][: s["nlocal"]] & gb) != 0
    for step, (ra, rb) in enumerate(zip(a, b)):
        for k in ra:
            if k == "T":   # T_a_i of atoms outside the group is never written by the reference (uninitialised storage)
                assert np.array_equal(ra[k][in_group], rb[k][in_group]), (step, k)
            else:
                assert np.array_equal(np.asarray(ra[k]), np.asarray(rb[k])), (step, k)
    if flags & 4 and flags & 3:
        assert a[-1]["Ee"] != a[0]["Ee"]          # the electronic energy moved


def test_atomic_heat_diffusion_from_gradient_bit_exact(refa):
    """heat_solve alone (flags 4): an energy gradient relaxes, the group's total energy is conserved to rounding
    except for what the one-sided clamp at zero adds"""
    s = H.make_system(3)
    drv = refa.atomic_fix_driver(s, H.atomic_fix_args(4, BETA, KAPPA, ["Ni"], inner_loops=4))
    fx = O.AtomicFix(s, O.Beta(path=BETA), O.Kappa(KAPPA), 4, inner_loops=4)
    E0 = drv.probe(6)[: s["nlocal"]] * (1.0 + 0.8 * np.sin(2 * np.pi * s["x"][: s["nlocal"], 0] / s["box"][0]))
    drv.set_energy(E0)
    fx.set_energy(E0)
    a = traj.run_atomic_fix_driver(drv, s, [None] * 4)
    b = traj.run_atomic_oracle(fx, s, [None] * 4, [58.71])
    for ra, rb in zip(a, b):
        for k in ("E", "T", "Ee", "Te", "array"):
            assert np.array_equal(np.asarray(ra[k]), np.asarray(rb[k])), k
    assert np.std(a[-1]["T"]) < np.std(a[0]["T"])
    assert abs(a[-1]["Ee"] - np.sum(E0)) < 1e-3 * np.sum(E0)


@pytest.mark.parametrize("name", ["atomic_caseA", "atomic_caseB_group"])
def test_atomic_oracle_matches_committed_golden_vectors(name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    s = traj.system_from_golden(g)
    gb = int(g["groupbit"])
    fx = O.AtomicFix(s, O.Beta(path=BETA), O.Kappa(KAPPA), int(g["flags"]), groupbit=gb, inner_loops=int(g["inner_loops"]),
                     dt=float(g["dt"]))
    if "E0" in g.files:
        fx.set_energy(g["E0"])
    recs = traj.run_atomic_oracle(fx, s, list(g["xi"]), [58.71])
    for k in KEYS:
        got = np.array([r[k] for r in recs])
        assert np.array_equal(got, g["out_" + k]), k


def test_reference_atomic_survives_atom_reordering(refa):
    """the stand-in's re-ordering (LAMMPS' spatial sort: copy_arrays along the permutation's cycles) is transparent to
    the reference fix, which carries E_a_i along (fix_eph_atomic.cpp:951-955) -- this pins the harness side of the
    migration tests of the product (test_atomic_emulated.py, test_zx_gpu_atomic.py)"""
    s = H.make_system(3)
    xi = [np.random.default_rng(40 + k).normal(size=(s["natoms"], 3)) for k in range(4)]
    mk = lambda system: refa.atomic_fix_driver(system, H.atomic_fix_args(7, BETA, KAPPA, ["Ni"], inner_loops=2))
    traj.assert_reordering_is_transparent(mk, s, xi, permute_after=2)


def test_atomic_three_elements_bit_exact(refa, tmp_path):
    """three atom types on three elements of both files (the smallest multi-element .kappa file the reference indexes in
    bounds: n_pairs = 4 >= 3, eph_kappa.h:69), element names given out of file order"""
    beta3 = str(H.write_beta_file(tmp_path / "synth3.beta", H.synthetic_knots(3, n_beta=5001, drho=0.01)))
    kappa3 = str(H.write_kappa_file(tmp_path / "synth3.kappa", H.synthetic_kappa(3, n_r=501, n_T=401, dT=2.5)))
    kr, ko = refa.Kappa(kappa3), O.Kappa(kappa3)
    assert (kr.n_elements, kr.n_pairs) == (3, 4) == (ko.n_elements, ko.n_pairs)
    for e in range(3):
        for kind in (0, 1, 2):
            assert np.array_equal(kr.table(kind, e), ko.table(kind, e)), (kind, e)
    for p in range(4):
        assert np.array_equal(kr.table(3, p), ko.table(3, p)), p
    s = H.make_system(3, ntypes=3, group_fraction=0.8)
    mass = [58.71, 58.93, 52.0]
    drv = refa.atomic_fix_driver(s, H.atomic_fix_args(7, beta3, kappa3, ["Cr", "Ni", "Co"], inner_loops=2, group="bit1"), mass=mass)
    fx = O.AtomicFix(s, O.Beta(path=beta3), ko, 7, groupbit=2, inner_loops=2, type_map_beta=[2, 0, 1], type_map_kappa=[2, 0, 1])
    xis = [np.random.default_rng(i).normal(size=(s["nlocal"], 3)) for i in range(3)]
    a = traj.run_atomic_fix_driver(drv, s, xis)
    b = traj.run_atomic_oracle(fx, s, xis, mass)
    in_group = (s["mask"][: s["nlocal"]] & 2) != 0
    for step, (ra, rb) in enumerate(zip(a, b)):
        for k in ra:
            if k == "T":
                assert np.array_equal(ra[k][in_group], rb[k][in_group]), (step, k)
            else:
                assert np.array_equal(np.asarray(ra[k]), np.asarray(rb[k])), (step, k)
