"""Generates the golden vectors under tests/golden/ from the UNMODIFIED reference
(compiled into oracle/_ref/libeph_ref.so) -- run in the development container:

    python tests/golden/make_golden.py

Inputs are stored next to the outputs so the fixtures do not depend on numpy's
generators staying bit-stable.  Ni_trunc.beta is a cut-down parametrisation for
these fixtures: the 1001 rho(r) knots and the first 201 beta(rho) knots
(rho <= 0.2 1/A^3, the live range of fcc Ni) of Data/Ni/Ni_PRB2019.beta.
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

NI = "/root/reference/Data/Ni/Ni_PRB2019.beta"


def write_ni_trunc(path, n_beta_keep=201):
    toks = open(NI).read().split("\n")
    head, body = toks[:5], " ".join(toks[5:]).split()
    n_rho, dr, n_beta, drho, rc = head[4].split()
    n_rho, n_beta = int(n_rho), int(n_beta)
    Z = body[0]
    rho = body[1:1 + n_rho]
    beta = body[1 + n_rho:1 + n_rho + n_beta_keep]
    with open(path, "w") as f:
        f.write("# cut-down fixture: rho(r) and the first %d beta(rho) knots of Ni_PRB2019\n#\n#\n" % n_beta_keep)
        f.write(head[3].strip() + "\n")
        f.write("%d %s %d %s %s\n" % (n_rho, dr, n_beta_keep, drho, rc))
        f.write(Z + "\n" + "\n".join(rho) + "\n" + "\n".join(beta) + "\n")


def pack(system, xis, recs, extra=None):
    d = dict(n=system["n"], x=system["x"], v=system["v"], type=system["type"], mask=system["mask"], tag=system["tag"],
             ghost_owner=system["ghost_owner"], nlocal=system["nlocal"], nghost=system["nghost"], box=system["box"],
             xi=np.array([np.zeros((system["nlocal"], 3)) if x is None else x for x in xis]))
    for k in ("f", "array", "T", "Ee", "Tmean", "w", "rho", "x", "v"):
        d["out_" + k] = np.array([r[k] for r in recs])
    d.update(extra or {})
    return d


def main():
    ni_trunc = os.path.join(HERE, "Ni_trunc.beta")
    write_ni_trunc(ni_trunc)
    rng = np.random.default_rng(20261017)

    # case A -- Examples/Example_1 geometry: 500 Ni atoms, flags 3, model 4, grid 1x1x1, T_e 300 (run.lmp:20)
    s = H.make_system(5)
    xis = [rng.normal(size=(s["nlocal"], 3)) for _ in range(3)]
    drv = R.fix_driver(s, H.fix_args(3, ni_trunc, ["Ni"], grid=(1, 1, 1)), dt=1e-4)
    recs = traj.run_fix_driver(drv, s, xis)
    np.savez_compressed(os.path.join(HERE, "caseA_example1.npz"), **pack(s, xis, recs, dict(flags=3, dt=1e-4)))

    # case B -- flags 7, FDM grid 4x3x2 from a grid file with a wall plane, a source, non-uniform kappa / C_e
    nx, ny, nz = 4, 3, 2
    ncell = nx * ny * nz
    box = [0, s["box"][0], 0, s["box"][1], 0, s["box"][2]]
    T0 = 300.0 + 50.0 * rng.random(ncell)
    kap = 0.1248 * (0.5 + rng.random(ncell))
    Ce = 3.5e-6 * (0.8 + 0.4 * rng.random(ncell))
    flag = np.ones(ncell, dtype=np.int64)
    flag[np.arange(ncell) % nx == 3] = 2      # zero-derivative wall plane i = 3
    flag[5] = 0                               # one constant cell
    S = np.zeros(ncell)
    S[1] = 1.0e-3
    gridfile = os.path.join(HERE, "caseB_grid.in")
    H.write_grid_file(gridfile, nx, ny, nz, box, T0, S, 1.0, Ce, kap, flag, 0, steps=2)
    xis = [rng.normal(size=(s["nlocal"], 3)) for _ in range(3)]
    cwd = os.getcwd()
    os.chdir(HERE)
    try:
        drv = R.fix_driver(s, H.fix_args(7, "Ni_trunc.beta", ["Ni"], T_infile="caseB_grid.in"), dt=1e-4)
        recs = traj.run_fix_driver(drv, s, xis)
    finally:
        os.chdir(cwd)
    np.savez_compressed(os.path.join(HERE, "caseB_grid.npz"), **pack(s, xis, recs, dict(flags=7, dt=1e-4)))

    # case C -- two elements, fix group = 60 % of the atoms (group "bit1"), synthetic tables, flags 7, grid 2x2x2
    s2 = H.make_system(4, ntypes=2, group_fraction=0.6, pos_seed=99, vel_seed=7)
    synth = os.path.join(HERE, "synth2.beta")
    H.write_beta_file(synth, H.synthetic_knots(2, n_beta=2001, drho=0.01))
    xis = [rng.normal(size=(s2["nlocal"], 3)) for _ in range(2)]
    drv = R.fix_driver(s2, H.fix_args(7, synth, ["Co", "Ni"], grid=(2, 2, 2), group="bit1"), dt=2e-4, mass=[58.93, 58.71])
    recs = traj.run_fix_driver(drv, s2, xis)
    np.savez_compressed(os.path.join(HERE, "caseC_alloy_group.npz"), **pack(s2, xis, recs, dict(flags=7, dt=2e-4)))
    print("golden vectors written to", HERE)


if __name__ == "__main__":
    main()
