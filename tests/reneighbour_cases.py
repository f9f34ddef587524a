"""Re-neighbouring with a changing number of ghosts (what LAMMPS does every few steps: ghosts and list rebuilt from the
current positions, atoms wrapped back into the box, per-atom arrays grown), product fix against the compiled reference
fix, both inside the LAMMPS stand-in.  Shared by the host-build tests and their `-m gpu` twins."""
import os

import numpy as np
import pytest

from eph_harness import harness as H
from eph_b200 import host

import traj
from conftest import GOLDEN

SCHEDULE = {2: 6.0, 4: 7.5, 5: 6.5}    # ghost shell / list cut-off before these steps: fewer, more, fewer ghosts
BETA = os.path.join(GOLDEN, "Ni_trunc.beta")
KAPPA = os.path.join(GOLDEN, "synth1.kappa")
TOL = 1e-10


def fix_case(style, extra, keywords=(), schedule=None):
    """style "eph" or "eph/coloured/exp": FixEPHB200 continues exactly like the reference (grid, filter state, energies).
    keywords: extra keyword pairs of the product fix (neigh device builds its list at r_c + neighbor->skin = 7 A: pass a
    schedule whose shells are 7 A)"""
    schedule = schedule or SCHEDULE
    from oracle import reference as R
    if not (R.available() and R.coloured_available()):
        pytest.skip("compiled reference not present")
    s = H.make_system(3, skin=2.0)
    xis = [np.random.default_rng(80 + k).normal(size=(s["nlocal"], 3)) for k in range(6)]
    ref_args = H.fix_args(7, BETA, ["Ni"], grid=(2, 2, 2), style=style, **extra)
    our_args = H.fix_args(7, BETA, ["Ni"], grid=(2, 2, 2), style=style + "/b200", extra=["rng", "mars"] + list(keywords), **extra)
    mk_ref = (lambda sy: R.coloured_fix_driver(sy, ref_args)) if "coloured" in style else (lambda sy: R.fix_driver(sy, ref_args))
    a = traj.run_with_reneighbouring(mk_ref, s, xis, schedule)
    b = traj.run_with_reneighbouring(lambda sy: host.FixDriver(sy, our_args), s, xis, schedule)
    assert len({r["nghost"] for r in a}) >= (3 if schedule is SCHEDULE else 1)
    traj.assert_same_trajectory(a, b, TOL)


def atomic_case(make_fix, comm):
    """the per-atom energies stay with their atoms: FixEPHAtomicB200 continues like the reference fix"""
    from oracle import reference as R
    if not R.atomic_available():
        pytest.skip("compiled reference not present")
    s = H.make_system(3, skin=2.0)
    xis = [np.random.default_rng(90 + k).normal(size=(s["nlocal"], 3)) for k in range(6)]
    ref_args = H.atomic_fix_args(7, BETA, KAPPA, ["Ni"], inner_loops=2)
    our_args = H.atomic_fix_args(7, BETA, KAPPA, ["Ni"], inner_loops=2, style="eph/atomic/b200") + ["rng", "mars", "comm", comm]
    a = traj.run_with_reneighbouring(lambda sy: R.atomic_fix_driver(sy, ref_args), s, xis, SCHEDULE)
    b = traj.run_with_reneighbouring(lambda sy: make_fix(sy, our_args), s, xis, SCHEDULE)
    assert len({r["nghost"] for r in a}) >= 3
    traj.assert_same_trajectory(a, b, TOL)


def adaptive_dt_case(kind, make_fix=None):
    """the cascade configuration runs with an adaptive time step (`Tests/MD_Run/run.lmp:94-97`): reset_dt
    at every change, product against the compiled reference (eta factor, grid dt and sub-stepping, memory-kernel zeta,
    ledger dt)"""
    from oracle import reference as R
    s = H.make_system(3, skin=2.0)
    xis = [np.random.default_rng(95 + k).normal(size=(s["nlocal"], 3)) for k in range(6)]
    dts = [5.5e-7, 5.5e-7, 2.0e-6, 1.0e-5, 1.0e-4, 1.0e-4]
    if kind == "atomic":
        if not R.atomic_available():
            pytest.skip("compiled reference not present")
        ref_args = H.atomic_fix_args(7, BETA, KAPPA, ["Ni"], inner_loops=2)
        our_args = H.atomic_fix_args(7, BETA, KAPPA, ["Ni"], inner_loops=2, style="eph/atomic/b200") + ["rng", "mars"]
        mk_ref, mk_our = (lambda sy: R.atomic_fix_driver(sy, ref_args)), (lambda sy: make_fix(sy, our_args))
    else:
        if not (R.available() and R.coloured_available()):
            pytest.skip("compiled reference not present")
        style, extra = ("eph/coloured/exp", dict(model="5e-4")) if kind == "coloured" else ("eph", {})
        ref_args = H.fix_args(7, BETA, ["Ni"], grid=(4, 4, 4), style=style, **extra)
        our_args = H.fix_args(7, BETA, ["Ni"], grid=(4, 4, 4), style=style + "/b200", extra=["rng", "mars"], **extra)
        mk_ref = (lambda sy: R.coloured_fix_driver(sy, ref_args)) if kind == "coloured" else (lambda sy: R.fix_driver(sy, ref_args))
        mk_our = lambda sy: host.FixDriver(sy, our_args)
    a = traj.run_with_reneighbouring(mk_ref, s, xis, {}, dts=dts)
    b = traj.run_with_reneighbouring(mk_our, s, xis, {}, dts=dts)
    traj.assert_same_trajectory(a, b, TOL, dts=dts)


def resident_case(cells=3, steps=7, every=3, sync=1, rng="mars"):
    """`integrate device`: x, v, f stay on the device between the hooks; FixEPHB200 in that mode continues exactly like the
    reference fix through re-neighbourings that happen, as in LAMMPS, between initial_integrate and post_force"""
    from oracle import reference as R
    if not R.available():
        pytest.skip("compiled reference not present")
    s = H.make_system(cells, skin=2.0)
    s["v"][0] = 300.0 * np.array([0.835115, 0.543981, 0.081652])     # one fast atom
    s["v"][s["nlocal"]:][s["ghost_owner"] == 0] = s["v"][0]
    if rng == "mars":
        xis = [np.random.default_rng(120 + k).normal(size=(s["nlocal"], 3)) for k in range(steps)]
    else:
        # the product's own counter-based stream (keyed on seed, step, atom tag): the reference is handed the same numbers.
        # Without injected Gaussians the fix starts the density pass of a step inside initial_integrate.
        from oracle import oracle as O
        xis = [O.xi_stream(12345, k + 1, s["tag"][: s["nlocal"]]) for k in range(steps)]
    ref_args = H.fix_args(7, BETA, ["Ni"], grid=(2, 2, 2))
    our_args = H.fix_args(7, BETA, ["Ni"], grid=(2, 2, 2), style="eph/b200",
                          extra=["rng", rng, "integrate", "device", "sync", sync])
    a = traj.run_in_lammps_order(lambda sy: R.fix_driver(sy, ref_args), s, xis, every)
    b = traj.run_in_lammps_order(lambda sy: host.FixDriver(sy, our_args, neigh_modify=(every, 0, False)), s, xis, every)
    if sync == 1:
        traj.assert_same_trajectory(a, b, TOL)
        return
    # `sync N`: LAMMPS' host f and v are current every N-th step only; x, the per-atom output and the energies every step
    for step, (ra, rb) in enumerate(zip(a, b), start=1):
        for key in ("x", "array") + (("v", "f") if step % sync == 0 else ()):
            assert H.error_metrics(rb[key], ra[key]) < TOL, (step, key)
        assert abs(ra["T"] - rb["T"]) <= TOL * abs(ra["T"])
