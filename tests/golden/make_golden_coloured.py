"""Generates the `fix eph/coloured/exp` golden vectors under tests/golden/ from the UNMODIFIED reference
(fix_eph_coloured_exp.cpp compiled into oracle/_ref/libeph_coloured_ref.so) -- run in the development container:

    python tests/golden/make_golden_coloured.py
"""
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

KEYS = ("f", "array", "T", "Ee", "Tmean", "w", "rho", "x", "v", "f_dis", "f_sto")


def main():
    beta = os.path.join(HERE, "Ni_trunc.beta")
    rng = np.random.default_rng(20261019)
    # 256 Ni atoms, fix group = 70 % of them, flags 7, tau0 = 5 dt, grid 2x2x2, 4 steps
    s = H.make_system(4, group_fraction=0.7)
    tau0 = 5.0e-4
    xis = [rng.normal(size=(s["nlocal"], 3)) for _ in range(4)]
    args = H.fix_args(7, beta, ["Ni"], model=repr(tau0), grid=(2, 2, 2), group="bit1", style="eph/coloured/exp")
    drv = R.coloured_fix_driver(s, args, dt=1e-4)
    recs = traj.run_fix_driver(drv, s, xis, vec3_probes=dict(f_dis=5, f_sto=6))
    d = dict(n=s["n"], x=s["x"], v=s["v"], type=s["type"], mask=s["mask"], tag=s["tag"], ghost_owner=s["ghost_owner"],
             nlocal=s["nlocal"], nghost=s["nghost"], box=s["box"], xi=np.array(xis), flags=7, dt=1e-4, tau0=tau0, groupbit=2)
    for k in KEYS:
        d["out_" + k] = np.array([r[k] for r in recs])
    np.savez_compressed(os.path.join(HERE, "coloured_case.npz"), **d)


if __name__ == "__main__":
    main()
