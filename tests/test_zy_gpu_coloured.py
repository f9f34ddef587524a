"""GPU tests of `fix eph/coloured/exp` on the device (eph_b200_set_colour, colour_filter_kernel behind the force pass),
written after this round's GPU budget was spent: their first run on a B200 is the driver's round-end run."""
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
    e_scale = 0.0
    for step, (a, b) in enumerate(zip(recs, refs)):
        for k in ("f", "array", "T", "w", "f_eph", "f_rng", "f_dis", "f_sto", "x", "v"):
            assert H.error_metrics(a[k], b[k]) < TOL, (step, k)
        e_scale += traj.energy_scale(b)
        assert abs(a["Ee"] - b["Ee"]) <= TOL * max(abs(b["Ee"]), e_scale, 1e-300), step
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
    got = np.array([r["Tmean"] for r in recs])
    assert np.all(np.abs(got - g["out_Tmean"]) <= TOL * np.abs(g["out_Tmean"]))
    scale = np.cumsum([traj.energy_scale(dict(array=g["out_array"][k], v=g["out_v"][k]), float(g["dt"])) for k in range(len(recs))])
    got = np.array([r["Ee"] for r in recs])
    assert np.all(np.abs(got - g["out_Ee"]) <= TOL * np.maximum(np.abs(g["out_Ee"]), scale))


def test_fix_coloured_b200_survives_atom_reordering(ni_trunc_beta):
    """LAMMPS re-orders the local atoms (spatial sort): the filter state migrates through copy_arrays and is registered
    again on the device, the trajectory continues unchanged"""
    s = H.make_system(3)
    xi = [np.random.default_rng(60 + k).normal(size=(s["natoms"], 3)) for k in range(4)]
    args = H.fix_args(7, ni_trunc_beta, ["Ni"], model="5e-4", grid=(2, 2, 2), style="eph/coloured/exp/b200", extra=["rng", "mars"])
    traj.assert_reordering_is_transparent(lambda system: host.FixDriver(system, args), s, xi, permute_after=2, tol=TOL)


@pytest.mark.parametrize("style,extra", [("eph", {}), ("eph/coloured/exp", dict(model="5e-4"))])
def test_fix_b200_through_reneighbouring_matches_reference(style, extra):
    """ghosts and list rebuilt from the current positions with a changing ghost count: FixEPHB200 (plain and coloured)
    continues exactly like the compiled reference fix"""
    import reneighbour_cases
    reneighbour_cases.fix_case(style, extra)


@pytest.mark.parametrize("kind", ["plain", "coloured"])
def test_fix_b200_adaptive_time_step_matches_reference(kind):
    """the cascade configuration's adaptive time step: reset_dt at every change (eta factor, grid dt and sub-step count,
    memory-kernel zeta), FixEPHB200 against the compiled reference fix"""
    import reneighbour_cases
    reneighbour_cases.adaptive_dt_case(kind)
