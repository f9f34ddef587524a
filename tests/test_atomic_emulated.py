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
    L = A.declare(C.CDLL(os.path.join(EMUL, "libeph_atomic_emul%s.so" % os.environ.get("EPH_EMUL_SUFFIX", ""))))
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
    from eph_harness import harness as H
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


def test_emulated_device_memspace(make_engine, kappa_tables):
    """memspace EPH_B200_DEVICE on the host build (device memory is host memory there): caller-owned type / mask / tag /
    list arrays are aliased, x, v are read and f is updated in place, without staging copies"""
    import numpy as np
    from eph_harness import harness as H
    from oracle import oracle as O
    import traj
    from test_engine_emulated import dev
    s = H.make_system(3)
    nl = s["nlocal"]
    eng = make_engine([0], [0], 7, inner_loops=2)
    from eph_b200 import host
    eng.set_tables_from(host.BetaTables(path=cases.BETA), kappa_tables)
    eng.set_dt(1e-4)
    keep = [dev(s["type"], np.int32), dev(s["mask"], np.int32), dev(s["tag"], np.int64), dev(s["ghost_owner"], np.int32),
            dev(s["offsets"], np.int64), dev(s["neigh"], np.int32)]
    eng.set_atoms(nl, s["nghost"], *keep[:4])
    eng.set_neighbors(*keep[4:])
    eng.init_energy(300.0)
    fx = O.AtomicFix(s, O.Beta(path=cases.BETA), O.Kappa(cases.KAPPA), 7, inner_loops=2)
    x, v = dev(s["x"].copy()), dev(s["v"].copy())
    for step in (1, 2):
        xi = np.random.default_rng(step).normal(size=(nl, 3))
        f = dev(np.zeros((nl, 3)))
        eng.post_force(x, v, f, dev(xi), step)
        Ee, Te = eng.end_of_step()
        fx.f[:] = 0.0
        fx.post_force(xi)
        fx.end_of_step()
        assert H.error_metrics(f.a, fx.f[:nl]) < cases.TOL
        assert H.error_metrics(eng.probe(6)[:nl], np.array(fx.ptr(6)[:nl])) < cases.TOL
        assert abs(Ee - fx.Ee()) <= cases.TOL * fx.Ee() and abs(Te - fx.Te()) <= cases.TOL * fx.Te()


@pytest.mark.parametrize("seed", [31, 32])
def test_fuzz_atomic_configurations(seed, make_engine, kappa_tables, tmp_path):
    import numpy as np
    from eph_harness import harness as H
    rng = np.random.default_rng(seed)
    beta2 = str(H.write_beta_file(tmp_path / "synth2.beta", H.synthetic_knots(2, n_beta=5001, drho=0.01)))
    for it in range(10):
        flags = int(rng.choice([1, 2, 3, 4, 5, 6, 7, 7 | 8, 7 | 16, 7 | 32, 7 | 16 | 32]))
        loops = int(rng.integers(0, 4))
        gf = None if rng.random() < 0.5 else float(rng.uniform(0.2, 0.9))
        n = int(rng.integers(2, 5))
        if rng.random() < 0.4:
            cases.trajectory_case(make_engine, kappa_tables, flags, loops, gf, n=n, steps=2, ntypes=2, beta=beta2, names=("Ni", "Co"))
        else:
            cases.trajectory_case(make_engine, kappa_tables, flags, loops, gf, n=n, steps=int(rng.integers(1, 4)))


@pytest.mark.parametrize("comm", ["device", "lammps"])
def test_emulated_fix_survives_atom_reordering(make_fix, comm):
    cases.reordering_case(make_fix, comm)


@pytest.mark.parametrize("comm", ["device", "lammps"])
def test_emulated_fix_through_reneighbouring_matches_reference(make_fix, comm):
    import reneighbour_cases
    reneighbour_cases.atomic_case(make_fix, comm)


def test_emulated_three_elements_in_both_files(make_engine, tmp_path):
    """three atom types on three elements of the .beta AND of the .kappa file: per-element locality densities, E(T) and
    K(T) tables (three elements give n_pairs = 4 >= 3, the smallest multi-element file the reference indexes in bounds,
    eph_kappa.h:69); a two-element .kappa file (n_pairs = 1) is refused"""
    from eph_harness import harness as H
    from eph_b200 import lib
    beta3 = str(H.write_beta_file(tmp_path / "synth3.beta", H.synthetic_knots(3, n_beta=5001, drho=0.01)))
    kappa3 = str(H.write_kappa_file(tmp_path / "synth3.kappa", H.synthetic_kappa(3, n_r=501, n_T=401, dT=2.5)))
    kt = A.KappaTables(kappa3)
    assert (kt.n_elements, kt.n_pairs) == (3, 4)
    cases.trajectory_case(make_engine, kt, 7, 2, 0.8, ntypes=3, beta=beta3, names=("Ni", "Co", "Cr"), kappa=kappa3, tk=[2, 0, 1])
    kappa2 = str(H.write_kappa_file(tmp_path / "synth2.kappa", H.synthetic_kappa(2, n_r=501, n_T=401, dT=2.5)))
    eng = make_engine([0, 1], [0, 1], 7)
    with pytest.raises(lib.EphError, match="indexed by element"):
        eng.set_tables_from(host_beta(beta3), A.KappaTables(kappa2))


def host_beta(path):
    from eph_b200 import host
    return host.BetaTables(path=path)


def test_emulated_call_order_errors_are_reported(make_engine, kappa_tables):
    """misuse of the C ABI comes back as an error code with a message, never as a crash or a silent no-op"""
    import numpy as np
    from eph_harness import harness as H
    from eph_b200 import host, lib
    s = H.make_system(2)
    nl = s["nlocal"]
    x, v, f = s["x"], s["v"], np.zeros((nl, 3))
    eng = make_engine([0], [0], 7, inner_loops=1)
    with pytest.raises(lib.EphError, match="tables not set"):
        eng.post_force(x, v, f, None, 1)
    eng.set_tables_from(host.BetaTables(path=cases.BETA), kappa_tables)
    with pytest.raises(lib.EphError, match="set_dt"):
        eng.post_force(x, v, f, None, 1)
    eng.set_dt(1e-4)
    with pytest.raises(lib.EphError, match="set_atoms"):
        eng.post_force(x, v, f, None, 1)
    ints = lambda a: np.ascontiguousarray(a, dtype=np.int32)
    with pytest.raises(lib.EphError, match="type"):
        eng.set_atoms(nl, s["nghost"], ints(s["type"]) + 5, ints(s["mask"]), np.ascontiguousarray(s["tag"], dtype=np.int64), ints(s["ghost_owner"]))
    with pytest.raises(lib.EphError, match="owner"):
        eng.set_atoms(nl, s["nghost"], ints(s["type"]), ints(s["mask"]), np.ascontiguousarray(s["tag"], dtype=np.int64), ints(s["ghost_owner"]) + nl)
    eng.set_atoms(nl, s["nghost"], ints(s["type"]), ints(s["mask"]), np.ascontiguousarray(s["tag"], dtype=np.int64), ints(s["ghost_owner"]))
    with pytest.raises(lib.EphError, match="set_neighbors"):
        eng.post_force(x, v, f, None, 1)
    bad = ints(s["neigh"]).copy()
    bad[3] = nl + s["nghost"] + 7
    with pytest.raises(lib.EphError, match="out of range"):
        eng.set_neighbors(np.ascontiguousarray(s["offsets"], dtype=np.int64), bad)
    eng.set_neighbors(np.ascontiguousarray(s["offsets"], dtype=np.int64), ints(s["neigh"]))
    eng.init_energy(300.0)
    with pytest.raises(lib.EphError, match="no post_force"):
        eng.end_of_step()                       # heat diffusion needs the positions of a post_force
    with pytest.raises(lib.EphError, match="without post_force_begin"):
        eng._check(eng.lib.eph_b200_atomic_post_force_mid(eng.h))
    eng.post_force(x, v, f, None, 1)
    Ee, Te = eng.end_of_step()
    assert Ee > 0 and 250 < Te < 350
    # the phase-split calls belong to the external-transport mode, the one-call ones to the internal one
    eng._check(eng.lib.eph_b200_atomic_set_comm_mode(eng.h, 1))
    with pytest.raises(lib.EphError, match="external comm mode"):
        eng.post_force(x, v, f, None, 2)
    with pytest.raises(lib.EphError, match="external comm mode"):
        eng.end_of_step()
    with pytest.raises(lib.EphError, match="unknown probe"):
        eng.probe(42)


def test_emulated_fix_adaptive_time_step_matches_reference(make_fix):
    import reneighbour_cases
    reneighbour_cases.adaptive_dt_case("atomic", make_fix)


def test_emulated_empty_and_ragged_inputs(make_engine, kappa_tables):
    cases.ragged_case(make_engine, kappa_tables)
