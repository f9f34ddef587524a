"""Shared driver of short `fix eph` trajectories for golden generation and parity tests.

One step = what LAMMPS' Verlet loop does around the fix with a
zero pair force: f := 0, initial_integrate, ghost x/v refresh, post_force,
final_integrate, ghost v refresh, end_of_step."""
import numpy as np


def energy_scale(rec, dt=1e-4):
    """Size of the terms of one step's energy transfer, sum_i (|f_EPH_i . v_i| + |f_RNG_i . v_i|) dt.  The cumulative energy
    `Ee` is a cancelling sum of such terms over atoms and steps (friction cools, the random force heats), so -- like the
    forces, SURVEY.md 8c -- it is compared against the size of its terms and not against what is left of them.
    rec: a record with the per-atom output `array` (columns 2..4 f_EPH, 5..7 f_RNG) and the local velocities `v`."""
    arr, v = np.asarray(rec["array"]), np.asarray(rec["v"])
    return float(np.abs(arr[:, 2:5] * v).sum() + np.abs(arr[:, 5:8] * v).sum()) * dt


class GhostSync:
    def __init__(self, system):
        self.nl = system["nlocal"]
        self.owner = np.asarray(system["ghost_owner"])
        x = np.asarray(system["x"])
        self.shift = x[self.nl:] - x[self.owner]

    def __call__(self, x, v):
        x[self.nl:] = x[self.owner] + self.shift
        v[self.nl:] = v[self.owner]


def run_fix_driver(drv, system, xis, dt=None, vec3_probes=None):
    """drv: eph_b200.host.FixDriver (reference or product).  xis: list of per-step xi [nlocal][3] or None.
    vec3_probes: {record key: probe id} of additional [nlocal][3] probes (fix eph/coloured/exp: f_dis 5, f_sto 6).
    Returns per-step records."""
    sync = GhostSync(system)
    out = []
    for step, xi in enumerate(xis, start=1):
        drv.set_step(step)
        x, v, f = drv.xvf()
        f[:] = 0.0
        drv.update(f=f)
        drv.initial_integrate()
        x, v, f = drv.xvf()
        sync(x, v)
        drv.update(x=x, v=v)
        if xi is not None:
            drv.set_xi(xi)
        drv.post_force()
        drv.final_integrate()
        x, v, f = drv.xvf()
        sync(x, v)
        drv.update(v=v)
        drv.end_of_step()
        x, v, f = drv.xvf()
        out.append(dict(x=x[: sync.nl].copy(), v=v[: sync.nl].copy(), f=f[: sync.nl].copy(), array=drv.array().copy(),
                        T=drv.grid_T().copy(), Ee=drv.compute_vector(0), Tmean=drv.compute_vector(1),
                        w=drv.probe(1).copy(), rho=drv.probe(0).copy()))
        for key, which in (vec3_probes or {}).items():
            out[-1][key] = drv.probe(which, vec3=True).copy()
    return out


def system_from_golden(g):
    """Rebuild the harness system dict (incl. the neighbour list) from a golden file's stored inputs."""
    from eph_harness import harness as H
    nl, ng = int(g["nlocal"]), int(g["nghost"])
    x = np.ascontiguousarray(g["x"])
    offsets, neigh = H.neighbor_list(x, nl, 7.0)
    return dict(n=int(g["n"]), natoms=nl, box=np.asarray(g["box"]), nlocal=nl, nghost=ng, x=x, v=np.ascontiguousarray(g["v"]),
                f=np.zeros_like(x), type=np.ascontiguousarray(g["type"]), mask=np.ascontiguousarray(g["mask"]),
                tag=np.ascontiguousarray(g["tag"]), ghost_owner=np.ascontiguousarray(g["ghost_owner"]), offsets=offsets,
                neigh=neigh, ntypes=int(np.max(g["type"])))


def run_oracle(fix, system, xis, mass):
    """fix: oracle.oracle.Fix.  Same step sequence as run_fix_driver."""
    sync = GhostSync(system)
    nl = sync.nl
    out = []
    for xi in xis:
        fix.f[:] = 0.0
        fix.initial_integrate(mass)
        sync(fix.x, fix.v)
        fix.post_force(xi)
        fix.final_integrate(mass)
        sync(fix.x, fix.v)
        fix.end_of_step()
        out.append(dict(x=fix.x[:nl].copy(), v=fix.v[:nl].copy(), f=fix.f[:nl].copy(), array=np.array(fix.ptr(5)),
                        T=np.array(fix.fdm.field(0)), Ee=fix.Ee(), Tmean=fix.fdm.T_total(), w=np.array(fix.ptr(1)),
                        rho=np.array(fix.ptr(0)), f_eph=np.array(fix.ptr(3)), f_rng=np.array(fix.ptr(4)),
                        f_dis=np.array(fix.ptr(6)), f_sto=np.array(fix.ptr(7))))
    return out


def run_engine(eng, system, xis, mass, dt, device=False, ftm2v=1.0 / 1.0364269e-4, coloured=False):
    """eng: eph_b200.lib.Engine with tables/grid/dt/atoms/neighbours set.  Drives the C ABI directly, with
    host (numpy) or device (torch) arrays.  `mass` is per type (1-based list without the leading slot)."""
    sync = GhostSync(system)
    nl = sync.nl
    m = np.concatenate([[0.0], np.atleast_1d(mass)])
    x = np.ascontiguousarray(system["x"], dtype=np.float64).copy()
    v = np.ascontiguousarray(system["v"], dtype=np.float64).copy()
    f = np.zeros((nl, 3))
    if device:
        import torch
        dev = torch.device("cuda", 0)
        owner = torch.as_tensor(sync.owner, device=dev, dtype=torch.long)
        shift = torch.as_tensor(sync.shift, device=dev)
        x, v, f = (torch.as_tensor(t, device=dev) for t in (x, v, f))

        def gsync():
            x[nl:] = x[owner] + shift
            v[nl:] = v[owner]
    else:
        def gsync():
            sync(x, v)
    out = []
    Ee = 0.0
    for step, xi in enumerate(xis, start=1):
        f[...] = 0.0
        eng.initial_integrate(x, v, f, m, dt, 0.5 * dt * ftm2v)
        gsync()
        xi_arg = xi
        if device and xi is not None:
            import torch
            xi_arg = torch.as_tensor(np.ascontiguousarray(xi), device=x.device)
        eng.post_force(x, v, f, xi_arg, step)
        eng.final_integrate(v, f, m, 0.5 * dt * ftm2v)
        gsync()
        Ee += eng.end_of_step(x, v)
        tonp = (lambda t: t.cpu().numpy()) if device else (lambda t: t.copy())
        out.append(dict(x=tonp(x[:nl]), v=tonp(v[:nl]), f=tonp(f[:nl]), array=eng.peratom(), T=eng.get_grid(0), Ee=Ee,
                        Tmean=eng.mean_T(), w=eng.probe(1), rho=eng.probe(0), f_eph=eng.probe(3), f_rng=eng.probe(4),
                        xi=eng.probe(2)))
        if coloured:
            out[-1]["f_dis"], out[-1]["f_sto"] = eng.colour_state()
    return out


# ---------------------------------------------------------------------------
# `fix eph/atomic`: the same Verlet sequence, with the per-atom electronic state in the records
# ---------------------------------------------------------------------------
def run_atomic_fix_driver(drv, system, xis):
    sync = GhostSync(system)
    nl = sync.nl
    out = []
    for step, xi in enumerate(xis, start=1):
        drv.set_step(step)
        x, v, f = drv.xvf()
        f[:] = 0.0
        drv.update(f=f)
        drv.initial_integrate()
        x, v, f = drv.xvf()
        sync(x, v)
        drv.update(x=x, v=v)
        if xi is not None:
            drv.set_xi(xi)
        drv.post_force()
        drv.final_integrate()
        x, v, f = drv.xvf()
        sync(x, v)
        drv.update(v=v)
        drv.end_of_step()
        x, v, f = drv.xvf()
        out.append(dict(x=x[:nl].copy(), v=v[:nl].copy(), f=f[:nl].copy(), array=drv.array().copy(), Ee=drv.compute_vector(0),
                        Te=drv.compute_vector(1), rho=drv.probe(0).copy(), w=drv.probe(1).copy(), f_eph=drv.probe(3).copy(),
                        f_rng=drv.probe(4).copy(), rho_a=drv.probe(5).copy(), E=drv.probe(6)[:nl].copy(),
                        dE=drv.probe(7).copy(), T=drv.probe(8).copy()))
    return out


def run_atomic_oracle(fix, system, xis, mass):
    """fix: oracle.oracle.AtomicFix"""
    sync = GhostSync(system)
    nl = sync.nl
    out = []
    for xi in xis:
        fix.f[:] = 0.0
        fix.initial_integrate(mass)
        sync(fix.x, fix.v)
        fix.post_force(xi)
        fix.final_integrate(mass)
        sync(fix.x, fix.v)
        fix.end_of_step()
        out.append(dict(x=fix.x[:nl].copy(), v=fix.v[:nl].copy(), f=fix.f[:nl].copy(), array=np.array(fix.ptr(9)), Ee=fix.Ee(),
                        Te=fix.Te(), rho=np.array(fix.ptr(0)), w=np.array(fix.ptr(1)), f_eph=np.array(fix.ptr(3)),
                        f_rng=np.array(fix.ptr(4)), rho_a=np.array(fix.ptr(5)), E=np.array(fix.ptr(6)[:nl]),
                        dE=np.array(fix.ptr(7)), T=np.array(fix.ptr(8))))
    return out


def run_atomic_engine(eng, system, xis, mass, dt, groupbit=1, noint=False, ftm2v=1.0 / 1.0364269e-4):
    """eng: eph_b200.atomic.AtomicEngine with tables/dt/atoms/neighbours/energies set.  Drives the C ABI with host
    arrays; the velocity-Verlet half steps are the host loops of FixEPHAtomicB200."""
    sync = GhostSync(system)
    nl = sync.nl
    x = np.ascontiguousarray(system["x"], dtype=np.float64).copy()
    v = np.ascontiguousarray(system["v"], dtype=np.float64).copy()
    f = np.zeros((nl, 3))
    g = (np.asarray(system["mask"][:nl]) & groupbit) != 0
    dtfm = (0.5 * dt * ftm2v / np.atleast_1d(mass)[np.asarray(system["type"][:nl]) - 1])[:, None]
    out = []
    for step, xi in enumerate(xis, start=1):
        f[...] = 0.0
        if not noint:
            v[:nl][g] += (dtfm * f)[g]
            x[:nl][g] += dt * v[:nl][g]
        sync(x, v)
        eng.post_force(x, v, f, None if xi is None else np.ascontiguousarray(xi), step)
        if not noint:
            v[:nl][g] += (dtfm * f)[g]
        sync(x, v)
        Ee, Te = eng.end_of_step()
        out.append(dict(x=x[:nl].copy(), v=v[:nl].copy(), f=f.copy(), array=eng.peratom(), Ee=Ee, Te=Te, rho=eng.probe(0),
                        w=eng.probe(1), f_eph=eng.probe(3), f_rng=eng.probe(4), rho_a=eng.probe(5), E=eng.probe(6)[:nl],
                        dE=eng.probe(7), T=eng.probe(8), xi=eng.probe(2)))
    return out


# ---------------------------------------------------------------------------
# atom re-ordering (LAMMPS' spatial sort): a fix must carry its per-atom state along through copy_arrays
# ---------------------------------------------------------------------------
def run_with_reordering(make_driver, system, xi_by_tag, permute_after=None, seed=5):
    """Runs len(xi_by_tag) Verlet steps of a fix living in the LAMMPS stand-in; after `permute_after` steps the local
    atoms are re-ordered at random (FixDriver.permute).  xi_by_tag[k]: Gaussians of step k+1 per atom TAG [natoms][3]
    (or None).  Returns per-step records with the per-atom quantities sorted by tag, so that runs with and without the
    re-ordering can be compared directly."""
    drv = make_driver(system)
    nl = system["nlocal"]
    tags = np.asarray(system["tag"][:nl]).copy()
    owner = np.asarray(system["ghost_owner"]).copy()
    x0 = np.asarray(system["x"])
    shift = x0[nl:] - x0[owner]
    out = []
    for k, xi_tag in enumerate(xi_by_tag):
        if permute_after is not None and k == permute_after:
            perm = np.random.default_rng(seed).permutation(nl).astype(np.int32)   # atom i moves to perm[i]
            drv.permute(perm)
            new_tags = np.empty_like(tags)
            new_tags[perm] = tags
            tags = new_tags
            owner = perm[owner]
        drv.set_step(k + 1)
        x, v, f = drv.xvf()
        f[:] = 0.0
        drv.update(f=f)
        drv.initial_integrate()
        x, v, f = drv.xvf()
        x[nl:] = x[owner] + shift
        v[nl:] = v[owner]
        drv.update(x=x, v=v)
        if xi_tag is not None:
            drv.set_xi(np.ascontiguousarray(xi_tag[tags - 1]))
        drv.post_force()
        drv.final_integrate()
        x, v, f = drv.xvf()
        v[nl:] = v[owner]
        drv.update(v=v)
        drv.end_of_step()
        x, v, f = drv.xvf()
        order = np.argsort(tags)
        out.append(dict(x=x[:nl][order].copy(), v=v[:nl][order].copy(), f=f[:nl][order].copy(), array=drv.array()[order].copy(),
                        Ee=drv.compute_vector(0), T=drv.compute_vector(1)))
    return out


def assert_reordering_is_transparent(make_driver, system, xi_by_tag, permute_after, tol=0.0):
    """the trajectory with a re-ordering after `permute_after` steps equals the one without (tol 0: bit for bit)"""
    from eph_harness import harness as H
    a = run_with_reordering(make_driver, system, xi_by_tag, None)
    b = run_with_reordering(make_driver, system, xi_by_tag, permute_after)
    for step, (ra, rb) in enumerate(zip(a, b), start=1):
        for key in ("x", "v", "f", "array"):
            if tol == 0.0:
                assert np.array_equal(ra[key], rb[key]), (step, key)
            else:
                assert H.error_metrics(rb[key], ra[key]) < tol, (step, key)
        for key in ("Ee", "T"):
            assert abs(ra[key] - rb[key]) <= max(tol, 1e-14) * abs(ra[key]), (step, key)
    # the state really mattered: without it the later steps would differ from a fresh start
    assert not np.array_equal(a[-1]["array"], a[0]["array"])


# ---------------------------------------------------------------------------
# re-neighbouring with a changing number of ghosts (what LAMMPS does every few steps)
# ---------------------------------------------------------------------------
def reneighbour(drv, system, shell):
    """Rebuilds ghosts and the full list from the CURRENT positions of the local atoms with a ghost shell / list cut-off
    of `shell`, and hands them to the fix in the stand-in the way LAMMPS does after re-neighbouring (the number of ghosts
    changes, per-atom arrays may have to grow).  Returns the new system dict."""
    from eph_harness import harness as H
    nl = system["nlocal"]
    x, v, f = drv.xvf()
    L = np.asarray(system["box"], dtype=np.float64)
    x[:nl] = np.mod(x[:nl], L)          # atoms that left the box are wrapped back first (Domain::remap)
    x[:nl][x[:nl] >= L] = 0.0
    xg, owner = H.ghosts(np.ascontiguousarray(x[:nl]), L, np.zeros(3), L.copy(), shell)
    owner = owner.astype(np.int32)
    s = dict(system)
    s["x"] = np.ascontiguousarray(np.concatenate([x[:nl], xg]))
    s["v"] = np.ascontiguousarray(np.concatenate([v[:nl], v[:nl][owner]]))
    s["f"] = np.ascontiguousarray(np.concatenate([f[:nl], np.zeros((len(xg), 3))]))
    for k in ("type", "mask", "tag"):
        loc = np.asarray(system[k][:nl])
        s[k] = np.ascontiguousarray(np.concatenate([loc, loc[owner]]))
    s["ghost_owner"], s["nghost"] = owner, len(xg)
    s["offsets"], s["neigh"] = H.neighbor_list(s["x"], nl, shell)
    drv.nghost = s["nghost"]
    drv._set_atoms(s)
    drv.set_neighbors(s["offsets"], s["neigh"])
    return s


def run_with_reneighbouring(make_driver, system, xis, schedule, dts=None, probes=False):
    """schedule: {step index: ghost shell} -- before that step the ghosts and the list are rebuilt with that shell;
    dts: time step of every step (LAMMPS' `fix dt/reset`: Fix::reset_dt is called whenever it changes);
    probes: also record rho_i (locals), w_i and the whole T_e grid of every step"""
    drv = make_driver(system)
    s = system
    out = []
    for k, xi in enumerate(xis):
        if dts is not None and (k == 0 or dts[k] != dts[k - 1]):
            drv.set_dt(dts[k])
        if k in schedule:
            s = reneighbour(drv, s, schedule[k])
        nl = s["nlocal"]
        sync = GhostSync(s)
        drv.set_step(k + 1)
        x, v, f = drv.xvf()
        f[:] = 0.0
        drv.update(f=f)
        drv.initial_integrate()
        x, v, f = drv.xvf()
        sync(x, v)
        drv.update(x=x, v=v)
        if xi is not None:
            drv.set_xi(xi)
        drv.post_force()
        drv.final_integrate()
        x, v, f = drv.xvf()
        sync(x, v)
        drv.update(v=v)
        drv.end_of_step()
        x, v, f = drv.xvf()
        out.append(dict(x=x[:nl].copy(), v=v[:nl].copy(), f=f[:nl].copy(), array=drv.array().copy(), Ee=drv.compute_vector(0),
                        T=drv.compute_vector(1), nghost=s["nghost"]))
        if probes:
            out[-1].update(rho=drv.probe(0)[:nl].copy(), w=drv.probe(1).copy(), grid=drv.grid_T().copy())
    return out


def assert_same_trajectory(a, b, tol, dts=None):
    from eph_harness import harness as H
    e_scale = 0.0
    for step, (ra, rb) in enumerate(zip(a, b), start=1):
        assert ra["nghost"] == rb["nghost"]
        for key in ("x", "v", "f", "array"):
            assert H.error_metrics(rb[key], ra[key]) < tol, (step, key)
        assert abs(ra["T"] - rb["T"]) <= tol * max(abs(ra["T"]), 1e-300), (step, "T")
        # the cumulative energy is a cancelling sum over atoms and steps (friction cools, the random force heats): like
        # the forces it is measured against the size of its terms; `fix eph` keeps f_EPH, f_RNG in columns 2..7
        if ra["array"].shape[1] == 8:
            dt = 1e-4 if dts is None else dts[step - 1]
            e_scale += energy_scale(ra, dt)
        assert abs(ra["Ee"] - rb["Ee"]) <= tol * max(abs(ra["Ee"]), e_scale, 1e-300), (step, "Ee")


def run_in_lammps_order(make_driver, system, xis, every, shell=7.0):
    """The Verlet loop in LAMMPS' own order -- initial_integrate, THEN (every `every`-th step) re-neighbouring from the host
    arrays, else the list ages and the ghosts are refreshed, post_force, final_integrate, end_of_step -- which is what a
    fix that keeps x, v, f on the device between the hooks (`integrate device`) has to cope with: LAMMPS migrates and
    re-orders atoms from its host arrays in the middle of a step."""
    drv = make_driver(system)     # (built with neigh_modify=(every, 0, False))
    s = system
    out = []
    for k, xi in enumerate(xis, start=1):
        nl = s["nlocal"]
        drv.set_step(k)
        drv.initial_integrate()     # first kick with last step's total force (zero before the first step), as in LAMMPS
        if k % every == 0:
            s = reneighbour(drv, s, shell)
        else:
            drv.neigh_tick()
            x, v, f = drv.xvf()
            GhostSync(s)(x, v)
            drv.update(x=x, v=v)
        x, v, f = drv.xvf()
        f[:] = 0.0
        drv.update(f=f)
        if xi is not None:
            drv.set_xi(xi)
        drv.post_force()
        drv.final_integrate()
        x, v, f = drv.xvf()
        GhostSync(s)(x, v)
        drv.update(v=v)
        drv.end_of_step()
        x, v, f = drv.xvf()
        out.append(dict(x=x[:nl].copy(), v=v[:nl].copy(), f=f[:nl].copy(), array=drv.array().copy(), Ee=drv.compute_vector(0),
                        T=drv.compute_vector(1), nghost=s["nghost"]))
    return out
