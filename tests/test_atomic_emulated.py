"""`fix eph/atomic` device path, checked WITHOUT a GPU: user-eph_b200/csrc/eph_atomic.cu -- the source nvcc compiles for
sm_100a -- is compiled for the host against a serial CUDA stand-in (tests/emul, test infrastructure) and driven through
the same C ABI, Python binding and FixEPHAtomicB200 host class as on the device, against the oracle at the 1e-10 bar.
This covers kernel arithmetic, indexing, flags, ghost handling and orchestration; sub-warp shuffles and real CUDA
semantics are what the `-m gpu` twins of these cases (tests/test_zz_gpu_late_additions.py) add on a B200."""
import ctypes as C
import os
import subprocess

import pytest

from eph_b200 import atomic as A

import atomic_cases as cases

EMUL = os.path.join(os.path.dirname(os.path.abspath(__file__)), "emul")


@pytest.fixture(scope="module")
def emul():
    subprocess.check_call(["make", "-C", EMUL, "libeph_atomic_emul.so"], stdout=subprocess.DEVNULL)
    # one self-contained library (engine + host class), loaded privately: see tests/emul/Makefile
    L = A.declare(C.CDLL(os.path.join(EMUL, "libeph_atomic_emul.so")))
    return L, L


@pytest.fixture(scope="module")
def make_engine(emul):
    return lambda tb, tk, flags, **kw: A.AtomicEngine(tb, tk, flags, lib=emul[0], **kw)


@pytest.fixture(scope="module")
def make_fix(emul):
    return lambda s, args: A.fix_driver(s, args, lib=emul[1])


@pytest.fixture(scope="module")
def kappa_tables():
    return A.KappaTables(cases.KAPPA)


@pytest.mark.parametrize("flags,loops,group_fraction", [(7, 0, None), (7, 3, None), (1, 0, None), (2, 0, None), (5, 2, None),
                                                        (6, 1, None), (7 + 16, 2, None), (7 + 32, 2, None), (7 + 8, 1, None),
                                                        (7, 2, 0.7), (4, 2, 0.5)])
def test_emulated_trajectory_matches_oracle(make_engine, kappa_tables, flags, loops, group_fraction):
    cases.trajectory_case(make_engine, kappa_tables, flags, loops, group_fraction)


def test_emulated_two_elements_in_the_beta_file(make_engine, kappa_tables, tmp_path):
    """two atom types on two .beta elements (g_ij != g_ji), both on the single .kappa element"""
    from eph_b200 import harness as H
    beta2 = str(H.write_beta_file(tmp_path / "synth2.beta", H.synthetic_knots(2, n_beta=5001, drho=0.01)))
    cases.trajectory_case(make_engine, kappa_tables, 7, 2, None, ntypes=2, beta=beta2, names=("Ni", "Co"))


def test_emulated_heat_diffusion_from_gradient(make_engine, kappa_tables):
    cases.gradient_case(make_engine, kappa_tables)


@pytest.mark.parametrize("name", ["atomic_caseA", "atomic_caseB_group"])
def test_emulated_engine_matches_committed_golden_vectors(make_engine, kappa_tables, name):
    cases.golden_engine_case(make_engine, kappa_tables, name)


@pytest.mark.parametrize("comm", ["device", "lammps"])
@pytest.mark.parametrize("name", ["atomic_caseA", "atomic_caseB_group"])
def test_emulated_fix_matches_committed_golden_vectors(make_fix, name, comm):
    cases.golden_fix_case(make_fix, name, comm)


def test_emulated_builtin_gaussian_stream(make_engine, kappa_tables):
    cases.philox_case(make_engine, kappa_tables)


def test_emulated_size_independent_properties(make_engine, kappa_tables):
    cases.properties_case(make_engine, kappa_tables, 4)
