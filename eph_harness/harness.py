"""The harness that plays LAMMPS for tests and benchmarks:
synthetic fcc lattices, Maxwell velocities, periodic ghost images, the full
neighbour list at r_c + skin, synthetic `.beta` / grid files, and a brick
decomposition for multi-GPU runs.  Everything is seeded and deterministic."""
import ctypes as C
import os

import numpy as np

from . import LIB, REPO_ROOT

KB = 8.617343e-5          # eV/K  (LAMMPS metal units, force->boltz)
MVV2E = 1.0364269e-4      # eV per amu (A/ps)^2
FTM2V = 1.0 / MVV2E
NI_MASS = 58.71
NI_A = 3.52

_h = None


def _hlib():
    global _h
    if _h is None:
        _h = C.CDLL(LIB)
        _h.eph_harness_ghosts.restype = C.c_longlong
    return _h


def shipped_beta(name):
    """Path of a shipped parametrisation (tests/golden/data/<name>.beta.gz, e.g. "Ni_PRB2019", "NiCoCrFe_PRB2019"),
    unpacked once per process into a temporary directory."""
    import gzip
    import tempfile
    cache = shipped_beta.__dict__.setdefault("cache", {})
    if name not in cache:
        src = os.path.join(REPO_ROOT, "tests", "golden", "data", name + ".beta.gz")
        dst = os.path.join(tempfile.mkdtemp(prefix="eph_beta_"), name + ".beta")
        with gzip.open(src, "rb") as f, open(dst, "wb") as g:
            g.write(f.read())
        cache[name] = dst
    return cache[name]


def fcc_positions(n, a=NI_A, sigma=0.05, seed=1234, sort_bin=3.5):
    """n^3 fcc unit cells (4 n^3 atoms) in a periodic box of side n*a, with N(0, sigma) displacements,
    wrapped into [0, L) and ordered by spatial bins like LAMMPS' `atom_modify sort`."""
    n = (n, n, n) if np.isscalar(n) else tuple(n)
    basis = np.array([[0, 0, 0], [0.5, 0.5, 0], [0.5, 0, 0.5], [0, 0.5, 0.5]], dtype=np.float64)
    ix, iy, iz = np.meshgrid(np.arange(n[0]), np.arange(n[1]), np.arange(n[2]), indexing="ij")
    cells = np.stack([ix.ravel(), iy.ravel(), iz.ravel()], axis=1).astype(np.float64)
    x = (cells[:, None, :] + basis[None, :, :]).reshape(-1, 3) * a
    L = np.array(n, dtype=np.float64) * a
    if sigma > 0:
        x = x + np.random.default_rng(seed).normal(0.0, sigma, x.shape)
    x = np.mod(x, L)
    x[x >= L] = 0.0
    if sort_bin:
        nb = np.maximum(1, np.floor(L / sort_bin).astype(np.int64))
        b = np.minimum((x / (L / nb)).astype(np.int64), nb - 1)
        key = (b[:, 2] * nb[1] + b[:, 1]) * nb[0] + b[:, 0]
        x = x[np.argsort(key, kind="stable")]
    return np.ascontiguousarray(x), L


def maxwell_velocities(natoms, T=600.0, mass=NI_MASS, seed=101):
    sigma = np.sqrt(KB * T / (mass * MVV2E))
    v = np.random.default_rng(seed).normal(0.0, sigma, (natoms, 3))
    v -= v.mean(axis=0)
    return np.ascontiguousarray(v)


def ghosts(xall, L, lo, hi, shell):
    """Ghost images around the brick [lo,hi): positions and the index (into xall) of each ghost's owner."""
    h = _hlib()
    xall = np.ascontiguousarray(xall, dtype=np.float64)
    L, lo, hi = (np.ascontiguousarray(t, dtype=np.float64) for t in (L, lo, hi))
    args = (C.c_longlong(len(xall)), C.c_void_p(xall.ctypes.data), C.c_void_p(L.ctypes.data), C.c_void_p(lo.ctypes.data),
            C.c_void_p(hi.ctypes.data), C.c_double(shell))
    n = h.eph_harness_ghosts(*args, None, None, None)
    xg = np.empty((n, 3), dtype=np.float64)
    owner = np.empty(n, dtype=np.int64)
    h.eph_harness_ghosts(*args, C.c_void_p(xg.ctypes.data), C.c_void_p(owner.ctypes.data), None)
    return xg, owner


def neighbor_list(x, nlocal, cut):
    """Full neighbour list (CSR) of atoms [0,nlocal) over all atoms in x within `cut`."""
    h = _hlib()
    x = np.ascontiguousarray(x, dtype=np.float64)
    nt = len(x)
    num = np.zeros(max(nlocal, 1), dtype=np.int32)
    h.eph_harness_neighbors(C.c_longlong(nlocal), C.c_longlong(nt), C.c_void_p(x.ctypes.data), C.c_double(cut),
                            C.c_void_p(num.ctypes.data), None, None)
    offsets = np.zeros(nlocal + 1, dtype=np.int64)
    np.cumsum(num[:nlocal], out=offsets[1:])
    flat = np.empty(max(int(offsets[-1]), 1), dtype=np.int32)
    h.eph_harness_neighbors(C.c_longlong(nlocal), C.c_longlong(nt), C.c_void_p(x.ctypes.data), C.c_double(cut),
                            C.c_void_p(num.ctypes.data), C.c_void_p(offsets.ctypes.data), C.c_void_p(flat.ctypes.data))
    return offsets, flat[: int(offsets[-1])]


def make_system(n, a=NI_A, sigma=0.05, T=600.0, r_cut=5.0, skin=2.0, ntypes=1, type_seed=7, group_fraction=None,
                pos_seed=1234, vel_seed=101, brick=None, mass=NI_MASS, with_list=True):
    """A periodic fcc system as one rank of a LAMMPS run would see it.

    brick = (rank, (px,py,pz)) cuts the rank's sub-domain out of the global box (spatial decomposition);
    None = the whole box on one rank.  Returns a dict with LAMMPS-layout arrays (locals first, then ghosts)."""
    xall, L = fcc_positions(n, a, sigma, pos_seed)
    natoms = len(xall)
    vall = maxwell_velocities(natoms, T, mass, vel_seed)
    tags = np.arange(1, natoms + 1, dtype=np.int64)
    if ntypes > 1:
        types_all = np.random.default_rng(type_seed).integers(1, ntypes + 1, natoms).astype(np.int32)
    else:
        types_all = np.ones(natoms, dtype=np.int32)
    mask_all = np.ones(natoms, dtype=np.int32)
    if group_fraction is not None:  # bit 1 (value 2) marks the fix group "bit1"; bit 0 is "all"
        sel = np.random.default_rng(type_seed + 1).random(natoms) < group_fraction
        mask_all = np.where(sel, 3, 1).astype(np.int32)
    if brick is None:
        lo, hi = np.zeros(3), L.copy()
        rank, grid = 0, (1, 1, 1)
    else:
        rank, grid = brick
        grid = tuple(grid)
        c = np.array([rank % grid[0], (rank // grid[0]) % grid[1], rank // (grid[0] * grid[1])])
        lo = L * c / np.array(grid)
        hi = L * (c + 1) / np.array(grid)
    inside = np.all((xall >= lo) & (xall < hi), axis=1)
    loc = np.nonzero(inside)[0]
    xg, owner_g = ghosts(xall, L, lo, hi, r_cut + skin)
    nlocal, nghost = len(loc), len(xg)
    x = np.ascontiguousarray(np.concatenate([xall[loc], xg]))
    v = np.ascontiguousarray(np.concatenate([vall[loc], vall[owner_g]]))
    sel = np.concatenate([loc, owner_g])
    glob2loc = np.full(natoms, -1, dtype=np.int64)
    glob2loc[loc] = np.arange(nlocal)
    if with_list:
        offsets, neigh = neighbor_list(x, nlocal, r_cut + skin)
    else:   # the caller builds the list on the device (eph_b200_build_neighbors)
        offsets, neigh = np.zeros(nlocal + 1, dtype=np.int64), np.zeros(1, dtype=np.int32)
    return dict(n=n, natoms=natoms, box=L, lo=lo, hi=hi, nlocal=nlocal, nghost=nghost, ntypes=ntypes, x=x, v=v,
                f=np.zeros_like(x), type=np.ascontiguousarray(types_all[sel]), mask=np.ascontiguousarray(mask_all[sel]),
                tag=np.ascontiguousarray(tags[sel]), ghost_owner=glob2loc[owner_g].astype(np.int32),
                ghost_owner_global=owner_g, local_global=loc, offsets=offsets, neigh=neigh, rank=rank, grid=grid,
                r_cut=r_cut, skin=skin, mass=mass)


# ---------------------------------------------------------------------------
# synthetic parametrisation: smooth positive rho(r), beta(rho) built from + - * /
# only, so that the file text is bit-reproducible on any machine.
# ---------------------------------------------------------------------------
def synthetic_knots(n_elements=1, n_rho=1001, n_beta=50001, r_cutoff=5.0, drho=0.001):
    dr = r_cutoff / (n_rho - 1)
    r = np.arange(n_rho, dtype=np.float64) * dr
    rho_axis = np.arange(n_beta, dtype=np.float64) * drho
    rho_k = np.empty((n_elements, n_rho))
    beta_k = np.empty((n_elements, n_beta))
    for e in range(n_elements):
        amp = 0.35 + 0.05 * e
        r0 = 1.9 + 0.1 * e
        t = 1.0 - r / r_cutoff
        t = np.where(t > 0.0, t, 0.0)
        q = r / r0
        rho_k[e] = amp * (t * t) * (t * t) / (0.02 + q * q * q * q)   # steep core, smooth (1-r/rc)^4 tail
        b0 = 0.25 + 0.03 * e
        c = 0.8 + 0.1 * e
        beta_k[e] = b0 * rho_axis * (rho_axis + 0.5 * c) / (rho_axis * rho_axis + c * rho_axis + 0.05)
    return n_elements, n_rho, dr, n_beta, drho, r_cutoff, rho_k, beta_k


def write_beta_file(path, knots, names=None, Z=None):
    """Write knots in the `.beta` grammar (reference Doc/Beta/input.beta)."""
    n_el, n_rho, dr, n_beta, drho, rc, rho_k, beta_k = knots
    names = names or ["Ni", "Co", "Cr", "Fe", "Al", "Cu"][:n_el]
    Z = Z or [28, 27, 24, 26, 13, 29][:n_el]
    with open(path, "w") as f:
        f.write("# synthetic electronic density and coupling, written by eph_harness\n# rho(r) [1/A^3], beta(rho) [eV ps/A^2]\n#\n")
        f.write("%d %s\n" % (n_el, " ".join(names)))
        f.write("%d %.17g %d %.17g %.17g\n" % (n_rho, dr, n_beta, drho, rc))
        for e in range(n_el):
            f.write("%d\n" % Z[e])
            f.write("\n".join("%.17e" % v for v in rho_k[e]))
            f.write("\n")
            f.write("\n".join("%.17e" % v for v in beta_k[e]))
            f.write("\n")
    return path


def write_grid_file(path, nx, ny, nz, box, T_e, S_e, rho_e, C_e, kappa_e, flag, t_dyn, steps=1, parameter_file="NULL"):
    """Write an FDM grid file in the current grammar (reference eph_fdm.h:48-119)."""
    n = nx * ny * nz

    def fld(v, dt=np.float64):
        a = np.empty(n, dtype=dt)
        a[...] = v
        return a

    T_e, S_e, rho_e, C_e, kappa_e = (fld(v) for v in (T_e, S_e, rho_e, C_e, kappa_e))
    flag, t_dyn = fld(flag, np.int64), fld(t_dyn, np.int64)
    with open(path, "w") as f:
        f.write("# grid written by eph_harness\n#\n#\n")
        f.write("%d %d %d %d\n" % (nx, ny, nz, steps))
        f.write("%.17e %.17e\n%.17e %.17e\n%.17e %.17e\n" % tuple(box))
        f.write("%s\n" % parameter_file)
        for k in range(nz):
            for j in range(ny):
                for i in range(nx):
                    r = i + j * nx + k * nx * ny
                    f.write("%d %d %d %.17e %.17e %.17e %.17e %.17e %d %d\n" % (i, j, k, T_e[r], S_e[r], rho_e[r], C_e[r],
                                                                            kappa_e[r], flag[r], t_dyn[r]))
    return path


def write_parameter_file(path, dT, C_e_T, kappa_e_T):
    with open(path, "w") as f:
        f.write("# C_e(T) kappa_e(T) written by eph_harness\n#\n#\n")
        f.write("%d %.17g\n" % (len(C_e_T), dT))
        for c, k in zip(C_e_T, kappa_e_T):
            f.write("%.17e %.17e\n" % (c, k))
    return path


def fix_args(flags, beta_file, elements, model=4, seed=12345, rho_e=1.0, C_e=3.5e-6, kappa_e=0.1248, T_e=300.0,
             grid=(1, 1, 1), T_infile="NULL", T_freq=0, T_out="T_out", group="all", style="eph", extra=()):
    """The `fix ID group eph ...` argument vector (reference fix_eph.cpp:36-58)."""
    return (["fx", group, style, seed, flags, model, repr(rho_e), repr(C_e), repr(kappa_e), repr(T_e), grid[0], grid[1],
             grid[2], T_infile, T_freq, T_out, beta_file] + list(elements) + list(extra))


# ---------------------------------------------------------------------------
# `fix eph/atomic` (reference fix_eph_atomic.cpp, eph_kappa.h): synthetic `.kappa` parametrisation
# ---------------------------------------------------------------------------
def synthetic_kappa(n_elements=1, n_r=1001, r_cutoff=4.5, n_T=2001, dT=1.0):
    """Locality rho_a(r), heat capacity C(T) and conductivity K(T) knots from + - * / only (bit-reproducible text)."""
    dr = r_cutoff / (n_r - 1)
    r = np.arange(n_r, dtype=np.float64) * dr
    T = np.arange(n_T, dtype=np.float64) * dT
    n_pairs = (n_elements + 1) * (n_elements - 1) // 2 if n_elements > 1 else 1   # eph_kappa.h:76
    rho_k = np.empty((n_elements, n_r))
    C_k = np.empty((n_elements, n_T))
    K_k = np.empty((n_pairs, n_T))
    for e in range(n_elements):
        t = 1.0 - r / r_cutoff
        t = np.where(t > 0.0, t, 0.0)
        rho_k[e] = (1.0 + 0.1 * e) * t * t * t / (1.0 + r * r / 4.0)
        C_k[e] = (1.0e-5 + 1.0e-6 * e) * (1.0 + T / 1000.0)
    for p in range(n_pairs):
        K_k[p] = (0.02 + 0.002 * p) * (1.0 + T / 1000.0)
    return n_elements, n_r, dr, r_cutoff, n_T, dT, T[-1], rho_k, C_k, K_k


def write_kappa_file(path, knots, names=None, Z=None):
    """Write knots in the `.kappa` grammar the reference reads (eph_kappa.h:53-151)."""
    n_el, n_r, dr, rc, n_T, dT, T_max, rho_k, C_k, K_k = knots
    names = names or ["Ni", "Co", "Cr", "Fe", "Al", "Cu"][:n_el]
    Z = Z or [28, 27, 24, 26, 13, 29][:n_el]
    with open(path, "w") as f:
        f.write("# synthetic per-atom electronic properties, written by eph_harness\n# rho_a(r), C(T), K(T)\n#\n")
        f.write("%d %s\n" % (n_el, " ".join(names)))
        f.write("%d %.17g %.17g %d %.17g %.17g\n" % (n_r, dr, rc, n_T, dT, T_max))
        for e in range(n_el):
            f.write("%d\n" % Z[e])
            f.write("\n".join("%.17e" % v for v in rho_k[e]))
            f.write("\n")
            f.write("\n".join("%.17e" % v for v in C_k[e]))
            f.write("\n")
        for p in range(len(K_k)):
            f.write("\n".join("%.17e" % v for v in K_k[p]))
            f.write("\n")
    return path


def atomic_fix_args(flags, beta_file, kappa_file, elements, seed=12345, T_e=300.0, T_infile="NULL", inner_loops=0,
                    T_out="NULL", group="all", style="eph/atomic"):
    """The `fix ID group eph/atomic ...` argument vector (reference fix_eph_atomic.cpp:39-56)."""
    return ["fx", group, style, seed, flags, repr(T_e), T_infile, inner_loops, T_out, beta_file, kappa_file] + list(elements)


def error_metrics(got, ref, floor=0.0):
    """max |got-ref| / max |ref|  (forces are cancelling sums, so errors are scaled by the largest reference magnitude)"""
    got, ref = np.asarray(got), np.asarray(ref)
    scale = max(float(np.max(np.abs(ref))) if ref.size else 0.0, floor)
    if scale == 0.0:
        return float(np.max(np.abs(got))) if got.size else 0.0
    return float(np.max(np.abs(got - ref))) / scale
