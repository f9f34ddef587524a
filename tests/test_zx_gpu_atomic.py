"""GPU tests of `fix eph/atomic` on the device (csrc/eph_atomic.cu behind include/eph_b200_atomic.h), written after this
round's GPU budget was spent: their first run on a B200 is the driver's round-end run.  The same cases already pass on
the CPU against a host build of the same source (tests/test_atomic_emulated.py); these are the real thing: the sm_100a
kernels with 8 lanes per atom, through the C ABI and through FixEPHAtomicB200."""
import numpy as np  # noqa: F401
import pytest

from eph_b200 import atomic as A
from eph_harness import harness as H

import atomic_cases as cases

pytestmark = pytest.mark.gpu


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


@pytest.mark.parametrize("comm", ["device", "lammps"])
@pytest.mark.parametrize("name", ["atomic_caseA", "atomic_caseB_group"])
def test_fix_atomic_b200_matches_committed_golden_vectors(name, comm):
    cases.golden_fix_case(lambda s, args: A.fix_driver(s, args), name, comm)


def test_atomic_engine_builtin_gaussian_stream(make_engine, kappa_tables):
    cases.philox_case(make_engine, kappa_tables)


def test_atomic_engine_properties_at_scale(make_engine, kappa_tables):
    """108 000 atoms (no oracle run needed): momentum, ledger = work, diffusion conserves, ghosts equal owners"""
    cases.properties_case(make_engine, kappa_tables, 30)


def test_fix_atomic_b200_survives_atom_reordering():
    cases.reordering_case(lambda s, args: A.fix_driver(s, args))


@pytest.mark.parametrize("comm", ["device", "lammps"])
def test_fix_atomic_b200_through_reneighbouring_matches_reference(comm):
    import reneighbour_cases
    reneighbour_cases.atomic_case(lambda s, args: A.fix_driver(s, args), comm)


def test_atomic_engine_three_elements_in_both_files(make_engine, tmp_path):
    """three atom types on three elements of the .beta and of the .kappa file (n_pairs = 4 >= 3, eph_kappa.h:69)"""
    beta3 = str(H.write_beta_file(tmp_path / "synth3.beta", H.synthetic_knots(3, n_beta=5001, drho=0.01)))
    kappa3 = str(H.write_kappa_file(tmp_path / "synth3.kappa", H.synthetic_kappa(3, n_r=501, n_T=401, dT=2.5)))
    cases.trajectory_case(make_engine, A.KappaTables(kappa3), 7, 2, 0.8, ntypes=3, beta=beta3, names=("Ni", "Co", "Cr"),
                          kappa=kappa3, tk=[2, 0, 1])


def test_fix_atomic_b200_adaptive_time_step_matches_reference():
    import reneighbour_cases
    reneighbour_cases.adaptive_dt_case("atomic", lambda s, args: A.fix_driver(s, args))


def test_atomic_engine_empty_and_ragged_inputs(make_engine, kappa_tables):
    cases.ragged_case(make_engine, kappa_tables)
