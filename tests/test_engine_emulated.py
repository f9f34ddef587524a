"""The `fix eph` engine, checked WITHOUT a GPU: user-eph_b200/csrc/eph_b200.cu and its headers -- the sources nvcc
compiles for sm_100a, with only the launch syntax rewritten (tests/emul/cu2cpp.py) -- are compiled for the host against
a lock-step SIMT stand-in (tests/emul: every CUDA thread a fiber, real sub-warp shuffles, ballots, block barriers and
shared memory) and driven through the same C ABI, Python binding and FixEPHB200 host class as on the device.  The test
bodies are the ones of the `-m gpu` suite (tests/test_gpu_parity.py, test_zy_gpu_coloured.py,
test_zz_gpu_late_additions.py), called here with the emulated library swapped in, against the oracle at the 1e-10 bar.
The TMA stencil kernels run too (a tensor map kept in the clear, box loads as synchronous copies with zero fill outside the
tensor, mbarrier waits already satisfied).  Not covered here: asynchrony and memory ordering, stream overlap, anything
about speed.  Test infrastructure only: the product has no CPU fallback."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from eph_harness import harness as H
from eph_b200 import host, lib
from oracle import oracle as O

import test_gpu_parity as G
import test_zy_gpu_coloured as GC
import test_zz_gpu_late_additions as GZ
import traj

EMUL = os.path.join(os.path.dirname(os.path.abspath(__file__)), "emul")


@pytest.fixture(scope="module", autouse=True)
def emulated_engine():
    """swap the emulated library in for the product's two (engine + host side) while this module runs"""
    subprocess.check_call(["make", "-C", EMUL, "libeph_b200_emul.so"], stdout=subprocess.DEVNULL)
    L = C.CDLL(os.path.join(EMUL, "libeph_b200_emul%s.so" % os.environ.get("EPH_EMUL_SUFFIX", "")))   # private: no RTLD_GLOBAL, linked -Bsymbolic
    for name, (res, args) in lib.SYMBOLS.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    L.ephh_last_error.restype = C.c_char_p
    for n in ("ephh_beta_load", "ephh_beta_from_knots", "ephh_grid_load"):
        getattr(L, n).restype = C.c_void_p
    L.ephh_grid_tables.restype = C.c_double
    L.emul_tma_load_count.restype = C.c_longlong
    saved = (lib._lib, host._fix)
    lib._lib, host._fix = L, L
    yield L
    lib._lib, host._fix = saved


@pytest.mark.parametrize("flags", [1, 3, 7, 2 | 4, 7 | 16, 7 | 32])
def test_emulated_engine_matches_oracle_flags(sys500, synth_beta_1, flags):
    G.test_engine_matches_oracle_flags(sys500, synth_beta_1, flags, False)


@pytest.mark.parametrize("lanes", ["1", "2", "4", "8", "16"])
def test_emulated_engine_lane_widths(sys500, synth_beta_1, lanes, monkeypatch):
    """every sub-warp width of the sweeps: 32 / lanes atoms per warp, group masks of `lanes` lanes"""
    monkeypatch.setenv("EPH_B200_LANES", lanes)
    s = sys500
    xis = [np.random.default_rng(3).normal(size=(s["nlocal"], 3))]
    fx = O.Fix(s, O.Beta(path=synth_beta_1), O.FDM(2, 2, 2, G.box6(s), 300.0, 3.5e-6, 1.0, 0.1248), 7, dt=1e-4)
    refs = traj.run_oracle(fx, s, xis, [58.71])
    eng = G.make_engine(synth_beta_1, 7, (2, 2, 2), G.box6(s))
    G.attach(eng, s)
    G.compare(traj.run_engine(eng, s, xis, [58.71], 1e-4), refs, s["nlocal"])


@pytest.mark.parametrize("model", [1, 2])
@pytest.mark.parametrize("flags", [1, 2 | 4, 7])
def test_emulated_legacy_models(sys500, synth_beta_1, model, flags):
    G.test_engine_legacy_models_match_oracle(sys500, synth_beta_1, model, flags)


@pytest.mark.parametrize("model", [1, 2])
def test_emulated_legacy_models_multi_element_and_group(synth_beta_4, model):
    G.test_engine_legacy_models_multi_element_and_group(synth_beta_4, model)


def test_emulated_multi_element_and_group(synth_beta_4):
    G.test_engine_multi_element_and_group(synth_beta_4)


def test_emulated_rho_above_cutoff(tmp_path):
    G.test_rho_above_cutoff_gives_zero_coupling(tmp_path)


def test_emulated_builtin_gaussian_stream(sys500, synth_beta_1):
    G.test_builtin_gaussian_stream_matches_its_definition(sys500, synth_beta_1)


@pytest.mark.parametrize("name", ["caseA_example1", "caseB_grid", "caseC_alloy_group"])
def test_emulated_engine_golden_vectors(name, ni_trunc_beta):
    G.test_engine_matches_committed_golden_vectors(name, ni_trunc_beta)


@pytest.mark.parametrize("comm", ["device", "lammps"])
@pytest.mark.parametrize("name", ["caseA_example1", "caseB_grid", "caseC_alloy_group"])
def test_emulated_fix_golden_vectors(name, comm):
    G.test_fix_b200_matches_committed_golden_vectors(name, comm)


@pytest.mark.parametrize("shape,walls,constant", [((8, 1, 1), False, False), ((5, 4, 3), True, True), ((1, 1, 1), False, False),
                                                  ((12, 5, 6), True, True), ((16, 4, 5), False, False)])
def test_emulated_grid_solve(synth_beta_1, shape, walls, constant):
    G.test_grid_solve_matches_oracle(synth_beta_1, shape, walls, constant)


@pytest.mark.parametrize("shape", [(32, 16, 8), (16, 4, 5)])
def test_emulated_grid_uniform_parameters(synth_beta_1, shape, emulated_engine):
    before = emulated_engine.emul_tma_load_count()
    G.test_grid_uniform_fast_path_matches_oracle(synth_beta_1, shape)
    assert emulated_engine.emul_tma_load_count() > before     # the constant-coefficient TMA kernel did the work


@pytest.mark.parametrize("shape", [(6, 5, 4), (32, 5, 4)])
def test_emulated_grid_temperature_dependent_cells(synth_beta_1, tmp_path, shape):
    G.test_grid_temperature_dependent_cells(synth_beta_1, tmp_path, shape)


@pytest.mark.parametrize("skin", [None, 2.0])
def test_emulated_inner_list_invalidation_and_rebuild(synth_beta_1, skin):
    G.test_inner_list_invalidation_and_rebuild(synth_beta_1, skin)


def test_emulated_empty_and_ragged_inputs(synth_beta_1):
    G.test_empty_and_ragged_inputs(synth_beta_1)


def test_emulated_device_built_neighbor_list(synth_beta_1):
    """eph_b200_build_neighbors (cell binning, radix sort, count + fill passes) against the host list"""
    s = H.make_system(5, sigma=0.08)
    nl = s["nlocal"]
    eng = G.make_engine(synth_beta_1, 7, (2, 2, 2), G.box6(s))
    eng.set_atoms(nl, s["nghost"], np.ascontiguousarray(s["type"], dtype=np.int32), np.ascontiguousarray(s["mask"], dtype=np.int32),
                  np.ascontiguousarray(s["tag"], dtype=np.int64), np.ascontiguousarray(s["ghost_owner"], dtype=np.int32))
    eng.build_neighbors(s["x"], 7.0)
    off, ne = eng.get_neighbors()
    assert np.array_equal(off, s["offsets"])
    for i in range(0, nl, 17):
        assert np.array_equal(np.sort(ne[off[i]:off[i + 1]]), np.sort(s["neigh"][s["offsets"][i]:s["offsets"][i + 1]]))
    xi = [np.random.default_rng(71).normal(size=(nl, 3)) for _ in range(2)]
    fx = O.Fix(s, O.Beta(path=synth_beta_1), O.FDM(2, 2, 2, G.box6(s), 300.0, 3.5e-6, 1.0, 0.1248), 7, dt=1e-4)
    G.compare(traj.run_engine(eng, s, xi, [58.71], 1e-4), traj.run_oracle(fx, s, xi, [58.71]), nl)


# ---- the paths written after round 1's GPU budget was spent: their GPU tests, run here first ----
@pytest.mark.parametrize("flags,group_fraction", [(7, None), (3, None), (1, None), (2, None), (7 + 16, None), (7 + 32, None),
                                                  (7 + 16 + 32, None), (7, 0.6)])
def test_emulated_coloured_engine(ni_trunc_beta, flags, group_fraction):
    GC.test_coloured_engine_matches_oracle(ni_trunc_beta, flags, group_fraction)


def test_emulated_coloured_fix_golden_vectors():
    GC.test_fix_coloured_b200_matches_committed_golden_vectors()


def test_emulated_coloured_fix_survives_atom_reordering(ni_trunc_beta):
    GC.test_fix_coloured_b200_survives_atom_reordering(ni_trunc_beta)


def test_emulated_fix_peratom_cadence():
    GZ.test_fix_b200_peratom_cadence()


def _host_views(engs, nz, plane):
    """the engines' current T_e buffers as numpy views: on the host build device memory is host memory"""
    out = []
    for e in engs:
        e.synchronize()
        p = C.c_void_p()
        e._check(e.lib.eph_b200_grid_device_ptr(e.h, 0, C.byref(p)))
        out.append(np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_double)), shape=(nz, plane)))
    return out


@pytest.mark.parametrize("shape,world,kind", [((32, 6, 8), 2, "walls"), ((32, 6, 8), 4, "walls"), ((33, 9, 6), 3, "walls"),
                                              ((32, 16, 8), 2, "uniform"), ((16, 4, 4), 4, "general")])
def test_emulated_sharded_grid_solve(synth_beta_1, shape, world, kind, monkeypatch, emulated_engine):
    """the plane range of the three stencil kernels and the plan / sub-step / external-finish calls of the sharded solve"""
    monkeypatch.setattr(GZ, "_views", _host_views)
    before = emulated_engine.emul_tma_load_count()
    GZ.test_sharded_grid_solve_matches_replicated_solve_and_oracle(synth_beta_1, shape, world, kind)
    # even nx: the plane range of the two TMA kernels; odd nx (33, 9, 6): that of the plain kernel
    assert (emulated_engine.emul_tma_load_count() > before) == (shape[0] % 2 == 0)


class DevArr:
    """what the Python binding takes for a device tensor (memspace EPH_B200_DEVICE), wrapped around a numpy array: on the
    host build device memory is host memory, so this drives the pointer-aliasing branches of the C ABI (no staging
    copies, f updated in place, caller-owned type / mask / tag / list arrays) that GPU-resident callers use.  `.a` is
    the array itself, for the harness' own arithmetic."""
    is_cuda = True

    def __init__(self, a, dtype=None):
        self.a = np.ascontiguousarray(a, dtype=dtype)

    def data_ptr(self):
        return self.a.ctypes.data

    def is_contiguous(self):
        return True

    def __len__(self):
        return len(self.a)


def dev(a, dtype=None):
    return DevArr(a, dtype)


@pytest.mark.parametrize("flags", [7, 3, 7 | 16 | 32])
def test_emulated_engine_device_memspace(sys500, synth_beta_1, flags):
    s = sys500
    nl = s["nlocal"]
    rng = np.random.default_rng(41)
    xis = [rng.normal(size=(nl, 3)) for _ in range(3)]
    fx = O.Fix(s, O.Beta(path=synth_beta_1), O.FDM(3, 2, 2, G.box6(s), 300.0, 3.5e-6, 1.0, 0.1248), flags, dt=1e-4)
    refs = traj.run_oracle(fx, s, xis, [58.71])
    eng = G.make_engine(synth_beta_1, flags, (3, 2, 2), G.box6(s))
    keep = [dev(s["type"], np.int32), dev(s["mask"], np.int32), dev(s["tag"], np.int64), dev(s["ghost_owner"], np.int32),
            dev(s["offsets"], np.int64), dev(s["neigh"], np.int32)]
    eng.set_atoms(nl, s["nghost"], *keep[:4])
    eng.set_neighbors(*keep[4:])
    sync = traj.GhostSync(s)
    x, v, f = dev(s["x"].copy()), dev(s["v"].copy()), dev(np.zeros((nl, 3)))
    m = np.array([0.0, 58.71])
    Ee, e_scale = 0.0, 0.0
    for step, (xi, ref) in enumerate(zip(xis, refs), start=1):
        f.a[...] = 0.0
        eng.initial_integrate(x, v, f, m, 1e-4, 0.5 * 1e-4 * H.FTM2V)
        sync(x.a, v.a)
        eng.post_force(x, v, f, dev(xi), step)
        eng.final_integrate(v, f, m, 0.5 * 1e-4 * H.FTM2V)
        sync(x.a, v.a)
        Ee += eng.end_of_step(x, v)
        for key, got in (("x", x.a[:nl]), ("v", v.a[:nl]), ("f", f.a), ("T", eng.get_grid(0)), ("array", eng.peratom()), ("w", eng.probe(1))):
            assert H.error_metrics(np.asarray(got), ref[key]) < G.TOL, (step, key)
        e_scale += traj.energy_scale(ref)
        assert abs(Ee - ref["Ee"]) <= G.TOL * max(abs(ref["Ee"]), e_scale, 1e-300)


def test_emulated_fix_integrate_device_matches_reference(emulated_engine):
    """keyword `integrate device`: the velocity-Verlet half steps run on the device and x, v, f stay there between the hooks"""
    import reneighbour_cases
    reneighbour_cases.resident_case()
    reneighbour_cases.resident_case(sync=2)
    reneighbour_cases.resident_case(rng="philox")


def test_simt_stand_in_selftest():
    """the lock-step stand-in itself: closed-form results for block barriers with early-exiting threads, sub-warp
    shuffles with different trip counts per group, ballots, 3-D launches and dynamic shared memory -- and a collective
    whose mask names a lane that never arrives must be reported as a dead-lock, not pass"""
    subprocess.check_call(["make", "-C", EMUL, "selftest"], stdout=subprocess.DEVNULL)
    exe = os.path.join(EMUL, "selftest")
    ok = subprocess.run([exe, "ok"], capture_output=True, text=True, timeout=120)
    assert ok.returncode == 0 and "selftest ok" in ok.stdout, ok.stdout + ok.stderr
    bad = subprocess.run([exe, "deadlock"], capture_output=True, text=True, timeout=120)
    assert bad.returncode != 0 and "dead-lock" in bad.stderr and "not reached" not in bad.stdout
    # a kernel that misses a barrier answers differently when the fibers are resumed in another order: running the
    # suites under SIMT_ORDER=reverse / random is the host build's racecheck (all product kernels are order-independent)
    sums = {mode: subprocess.run([exe, "race"], capture_output=True, text=True, timeout=120,
                                 env=dict(os.environ, SIMT_ORDER=mode)).stdout for mode in ("forward", "reverse")}
    assert "race checksum" in sums["forward"] and sums["forward"] != sums["reverse"], sums


# ---- re-neighbouring with a changing ghost count, product against the compiled reference, both in the stand-in ----
@pytest.mark.parametrize("style,extra", [("eph", {}), ("eph/coloured/exp", dict(model="5e-4"))])
def test_emulated_fix_through_reneighbouring_matches_reference(style, extra):
    import reneighbour_cases
    reneighbour_cases.fix_case(style, extra)


@pytest.mark.parametrize("keywords", [("neigh", "device"), ("comm", "lammps"), ("neigh", "device", "comm", "lammps")])
def test_emulated_fix_keywords_through_reneighbouring(keywords):
    """the list built on the device from the positions and / or the ghost values through LAMMPS' forward comm, across
    re-neighbourings"""
    import reneighbour_cases
    reneighbour_cases.fix_case("eph", {}, keywords, schedule={2: 7.0, 4: 7.0})


def test_emulated_long_trajectory_list_state_machine(ni_trunc_beta):
    """200 velocity-Verlet steps of hot atoms under the fix's own forces, LAMMPS' list never rebuilt: the inner list is
    rebuilt several times as the atoms drift; late rebuilds stand for a shorter skin (skin - 2 D: the guard trips sooner)
    and only when that falls below 0.1 A does the engine stay on LAMMPS' list -- 20 fall-back steps here, 51 when a late
    rebuild had to carry the full inner skin -- and the forces, positions and grid follow the oracle to 1e-10 all the way"""
    s = H.make_system(4, T=6000.0)
    dt, nl = 1.5e-4, s["nlocal"]
    rng = np.random.default_rng(3)
    fx = O.Fix(s, O.Beta(path=ni_trunc_beta), O.FDM(2, 2, 2, G.box6(s), 300.0, 3.5e-6, 1.0, 0.1248), 7, dt=dt)
    eng = G.make_engine(ni_trunc_beta, 7, (2, 2, 2), G.box6(s), dt=dt)
    eng.set_skin(2.0, 0.4)
    G.attach(eng, s)
    sync = traj.GhostSync(s)
    x, v = s["x"].copy(), s["v"].copy()
    f = np.zeros((nl, 3))
    m, dtf = np.array([0.0, 58.71]), 0.5 * dt * H.FTM2V
    fx.f[:] = 0.0
    worst = 0.0
    for step in range(1, 201):
        xi = rng.normal(size=(nl, 3))
        eng.initial_integrate(x, v, f, m, dt, dtf)
        sync(x, v)
        f = np.zeros((nl, 3))
        eng.post_force(x, v, f, xi, step)
        eng.final_integrate(v, f, m, dtf)
        sync(x, v)
        eng.end_of_step(x, v)
        fx.initial_integrate([58.71])
        sync(fx.x, fx.v)
        fx.f[:] = 0.0
        fx.post_force(xi)
        fx.final_integrate([58.71])
        sync(fx.x, fx.v)
        fx.end_of_step()
        worst = max(worst, H.error_metrics(f, fx.f[:nl]), H.error_metrics(x[:nl], fx.x[:nl]),
                    H.error_metrics(eng.get_grid(0), fx.fdm.field(0)))
    st = eng.list_stats()
    assert worst < G.TOL, worst
    assert st["inner_builds"] >= 5 and 8 <= st["fallback_steps"] <= 35, st
    assert 0.8 < np.abs(x[:nl] - s["x"][:nl]).max() < 1.2


def test_emulated_long_trajectory_with_a_longer_list(ni_trunc_beta):
    """The same 170 hot steps with LAMMPS' list 0.4 A longer (fix keyword `extra_skin 0.4`: cut-off r_c + 2.4 A while LAMMPS
    still re-neighbours at half of ITS 2 A skin, i.e. before any atom has moved 1 A): the inner list can be rebuilt from
    the aged list for the whole life of the list -- the engine never gives up and stays on the full list -- and
    the results follow the oracle as before."""
    s = H.make_system(4, T=6000.0, skin=2.4)
    dt, nl = 1.5e-4, s["nlocal"]
    rng = np.random.default_rng(3)
    fx = O.Fix(s, O.Beta(path=ni_trunc_beta), O.FDM(2, 2, 2, G.box6(s), 300.0, 3.5e-6, 1.0, 0.1248), 7, dt=dt)
    eng = G.make_engine(ni_trunc_beta, 7, (2, 2, 2), G.box6(s), dt=dt)
    eng.set_skin(2.4, 0.4)      # what FixEPHB200::init() passes: neighbor->skin + extra_skin
    G.attach(eng, s)
    sync = traj.GhostSync(s)
    x, v = s["x"].copy(), s["v"].copy()
    f = np.zeros((nl, 3))
    m, dtf = np.array([0.0, 58.71]), 0.5 * dt * H.FTM2V
    fx.f[:] = 0.0
    worst, trips = 0.0, 0
    for step in range(1, 171):
        xi = rng.normal(size=(nl, 3))
        eng.initial_integrate(x, v, f, m, dt, dtf)
        sync(x, v)
        f = np.zeros((nl, 3))
        eng.post_force(x, v, f, xi, step)
        eng.final_integrate(v, f, m, dtf)
        sync(x, v)
        eng.end_of_step(x, v)
        fx.initial_integrate([58.71])
        sync(fx.x, fx.v)
        fx.f[:] = 0.0
        fx.post_force(xi)
        fx.final_integrate([58.71])
        sync(fx.x, fx.v)
        fx.end_of_step()
        worst = max(worst, H.error_metrics(f, fx.f[:nl]), H.error_metrics(x[:nl], fx.x[:nl]),
                    H.error_metrics(eng.get_grid(0), fx.fdm.field(0)))
    st = eng.list_stats()
    moved = np.abs(x[:nl] - s["x"][:nl]).max()
    assert worst < G.TOL, worst
    assert 0.7 < moved < 1.0, moved          # within the life of a LAMMPS list (re-neighbouring at 1 A)
    # every guard trip costs one step on the full list and is answered by a rebuild: no run of fall-back steps
    assert st["inner_builds"] >= 4 and st["fallback_steps"] <= st["inner_builds"], st


def test_emulated_fix_extra_skin_keyword():
    """`extra_skin 0.4`: the fix asks LAMMPS for a list 0.4 A longer and tells the engine so; forces, grid and energies are
    those of the reference on the standard list (r_c decides the pair set), across re-neighbourings; a negative value is
    refused"""
    import reneighbour_cases
    reneighbour_cases.fix_case("eph", {}, ("extra_skin", "0.4"), schedule={2: 7.4, 4: 7.4})
    s = H.make_system(2, skin=2.4)
    bad = H.fix_args(7, reneighbour_cases.BETA, ["Ni"], grid=(1, 1, 1), style="eph/b200", extra=["extra_skin", "-0.1"])
    with pytest.raises(Exception, match="extra_skin must be >= 0"):
        host.FixDriver(s, bad)


def test_emulated_fixes_with_32_bit_atom_tags(ni_trunc_beta, emulated_engine):
    """LAMMPS' default build (-DLAMMPS_SMALLBIG) has 32-bit atom tags: the host classes, compiled against a stand-in with
    `typedef int tagint`, widen the tags for the C ABI and give the same forces -- with the built-in Gaussian stream too,
    which is keyed on the tags"""
    from conftest import GOLDEN
    from eph_b200 import atomic as A
    subprocess.check_call(["make", "-C", EMUL, "TAG32=1", "SUFFIX=_tag32", "libeph_b200_emul_tag32.so", "libeph_atomic_emul_tag32.so"],
                          stdout=subprocess.DEVNULL)
    L32 = C.CDLL(os.path.join(EMUL, "libeph_b200_emul_tag32.so"))
    g = np.load(os.path.join(GOLDEN, "caseA_example1.npz"))
    s = traj.system_from_golden(g)
    s["natoms"] = s["nlocal"]
    s["tag"] = np.ascontiguousarray(np.random.default_rng(7).permutation(s["nlocal"])[np.r_[np.arange(s["nlocal"]), s["ghost_owner"]]] + 1)
    args = H.fix_args(3, ni_trunc_beta, ["Ni"], grid=(1, 1, 1), style="eph/b200")     # rng philox: xi from (seed, tag, step)
    xis = [None] * 3
    a = traj.run_fix_driver(host.FixDriver(s, args, lib=L32), s, xis)
    b = traj.run_fix_driver(host.FixDriver(s, args), s, xis)
    for ra, rb in zip(a, b):
        assert np.array_equal(ra["f"], rb["f"]) and np.abs(ra["f"]).max() > 0
    import atomic_cases as cases
    La32 = C.CDLL(os.path.join(EMUL, "libeph_atomic_emul_tag32.so"))
    La = C.CDLL(os.path.join(EMUL, "libeph_atomic_emul.so"))
    aargs = H.atomic_fix_args(7, cases.BETA, cases.KAPPA, ["Ni"], inner_loops=1, style="eph/atomic/b200")
    a = traj.run_atomic_fix_driver(A.fix_driver(s, aargs, lib=La32), s, xis)
    b = traj.run_atomic_fix_driver(A.fix_driver(s, aargs, lib=La), s, xis)
    for ra, rb in zip(a, b):
        assert np.array_equal(ra["f"], rb["f"]) and np.array_equal(ra["E"], rb["E"]) and np.abs(ra["f_rng"]).max() > 0


def test_emulated_fix_output_files_match_reference(ni_trunc_beta, tmp_path):
    """the files the fix writes: heat maps `T_out_%06d` every `freq` steps (fix_eph.cpp:397-399) and the restart file of
    post_run (fix_eph.cpp:1019-1021; `<T_infile>.restart` when the grid came from a file) -- product against the
    compiled reference, both inside the stand-in, each in its own working directory; the restart is then read back as
    T_infile of a continuation run"""
    from oracle import reference as R
    if not R.available():
        pytest.skip("compiled reference not present")
    s = H.make_system(3)
    xis = [np.random.default_rng(70 + k).normal(size=(s["nlocal"], 3)) for k in range(4)]
    nc = 4 * 3 * 2
    T0 = 300 + 50 * np.random.default_rng(10).random(nc)
    fl = np.ones(nc, dtype=np.int64)
    fl[::5] = 2
    fl[7] = 0
    out = {}
    for who in ("ref", "b200"):
        d = tmp_path / who
        d.mkdir()
        H.write_grid_file(d / "T.in", 4, 3, 2, G.box6(s), T0, 0.0, 1.0, 3.5e-6, 0.1248, fl, 0, steps=1)
        cwd = os.getcwd()
        os.chdir(d)
        try:
            style = "eph" if who == "ref" else "eph/b200"
            extra = [] if who == "ref" else ["rng", "mars"]
            args = H.fix_args(7, ni_trunc_beta, ["Ni"], T_infile="T.in", T_freq=2, T_out="T_out", style=style, extra=extra)
            drv = R.fix_driver(s, args) if who == "ref" else host.FixDriver(s, args)
            recs = traj.run_fix_driver(drv, s, xis)
            drv.post_run()
            out[who] = dict(files=sorted(os.listdir(d)), recs=recs,
                            content={f: open(d / f).read() for f in os.listdir(d) if f != "T.in"})
        finally:
            os.chdir(cwd)
    assert out["ref"]["files"] == out["b200"]["files"] == ["T.in", "T.in.restart", "T_out_000001", "T_out_000002"]
    for name, text in out["ref"]["content"].items():
        ours = out["b200"]["content"][name]
        if ours != text:   # %e formatting of values that agree to 1e-14 can differ in the last printed digit
            a = np.array([float(t) for t in text.split() if t[0].isdigit() or t[0] in "+-."], dtype=float)
            b = np.array([float(t) for t in ours.split() if t[0].isdigit() or t[0] in "+-."], dtype=float)
            assert a.shape == b.shape and np.allclose(a, b, rtol=2e-6, atol=0), name
        assert text.splitlines()[:7] == ours.splitlines()[:7], name           # headers identical
    # continuation from the product's restart file equals continuation from the reference's
    for who in ("ref", "b200"):
        cwd = os.getcwd()
        os.chdir(tmp_path / who)
        try:
            args = H.fix_args(7, ni_trunc_beta, ["Ni"], T_infile="T.in.restart", style="eph/b200", extra=["rng", "mars"])
            out[who + "2"] = traj.run_fix_driver(host.FixDriver(s, args), s, xis[:2])
        finally:
            os.chdir(cwd)
    for ra, rb in zip(out["ref2"], out["b2002"]):
        assert H.error_metrics(rb["T"], ra["T"]) < 1e-6 and H.error_metrics(rb["f"], ra["f"]) < 1e-6


@pytest.mark.parametrize("kind", ["plain", "coloured"])
def test_emulated_fix_adaptive_time_step_matches_reference(kind):
    import reneighbour_cases
    reneighbour_cases.adaptive_dt_case(kind)
