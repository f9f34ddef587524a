"""Generates the `fix eph/atomic` golden vectors under tests/golden/ from the UNMODIFIED reference
(fix_eph_atomic.cpp compiled into oracle/_ref/libeph_atomic_ref.so) -- run in the development container:

    python tests/golden/make_golden_atomic.py

synth1.kappa is a synthetic per-atom parametrisation (eph_harness.harness.synthetic_kappa) in the reference's `.kappa`
grammar; inputs are stored next to the outputs."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "user-eph_b200"))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from eph_harness import harness as H  # noqa: E402
from oracle import reference as R  # noqa: E402
import traj  # noqa: E402

KEYS = ("f", "array", "Ee", "Te", "rho", "w", "f_eph", "f_rng", "rho_a", "E", "dE", "x", "v")


def pack(system, xis, recs, extra):
    d = dict(n=system["n"], x=system["x"], v=system["v"], type=system["type"], mask=system["mask"], tag=system["tag"],
             ghost_owner=system["ghost_owner"], nlocal=system["nlocal"], nghost=system["nghost"], box=system["box"],
             xi=np.array([np.zeros((system["nlocal"], 3)) if x is None else x for x in xis]))
    for k in KEYS:
        d["out_" + k] = np.array([r[k] for r in recs])
    d.update(extra)
    return d


def main():
    kappa = os.path.join(HERE, "synth1.kappa")
    H.write_kappa_file(kappa, H.synthetic_kappa(1, n_r=501, r_cutoff=4.5, n_T=401, dT=2.5))
    beta = os.path.join(HERE, "Ni_trunc.beta")
    rng = np.random.default_rng(20261018)

    # case A -- 256 Ni atoms, friction + random + heat diffusion (flags 7), 2 inner loops, all atoms in the group
    s = H.make_system(4)
    xis = [rng.normal(size=(s["nlocal"], 3)) for _ in range(3)]
    drv = R.atomic_fix_driver(s, H.atomic_fix_args(7, beta, kappa, ["Ni"], inner_loops=2), dt=1e-4)
    recs = traj.run_atomic_fix_driver(drv, s, xis)
    np.savez_compressed(os.path.join(HERE, "atomic_caseA.npz"), **pack(s, xis, recs, dict(flags=7, dt=1e-4, inner_loops=2, groupbit=1)))

    # case B -- fix group = 70 % of the atoms (group "bit1"), an initial energy gradient along x, flags 7, 1 loop
    s = H.make_system(4, group_fraction=0.7)
    xis = [rng.normal(size=(s["nlocal"], 3)) for _ in range(3)]
    drv = R.atomic_fix_driver(s, H.atomic_fix_args(7, beta, kappa, ["Ni"], inner_loops=0, group="bit1"), dt=1e-4)
    E0 = drv.probe(6)[: s["nlocal"]] * (1.0 + 0.5 * s["x"][: s["nlocal"], 0] / s["box"][0])
    drv.set_energy(E0)
    recs = traj.run_atomic_fix_driver(drv, s, xis)
    np.savez_compressed(os.path.join(HERE, "atomic_caseB_group.npz"),
                        **pack(s, xis, recs, dict(flags=7, dt=1e-4, inner_loops=0, groupbit=2, E0=E0)))


if __name__ == "__main__":
    main()
