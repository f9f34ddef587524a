"""TEST INFRASTRUCTURE: ctypes binding of the C restatement (oracle/liboracle.so).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg import this."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


def build():
    subprocess.check_call(["make", "-C", HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    subprocess.check_call(["make", "-C", HERE, "ref"], stdout=subprocess.DEVNULL)


def load():
    global _lib
    if _lib is None:
        path = os.path.join(HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        for n in ("orc_beta_load", "orc_beta_from_knots", "orc_fdm_new", "orc_fdm_from_file", "orc_fix_new",
                  "orc_beta_table", "orc_fdm_field", "orc_fdm_flags", "orc_fdm_tdyn", "orc_fix_ptr", "orc_kappa_load",
                  "orc_kappa_table", "orc_afix_new", "orc_afix_ptr"):
            getattr(L, n).restype = C.c_void_p
        for n in ("orc_spline_eval", "orc_linear_eval", "orc_linear_reverse", "orc_beta_rho_r_sq", "orc_beta_alpha",
                  "orc_beta_beta", "orc_fdm_get_T", "orc_fdm_T_total", "orc_fix_Ee", "orc_afix_Ee", "orc_afix_Te"):
            getattr(L, n).restype = C.c_double
        L.orc_fdm_index.restype = C.c_size_t
        L.orc_sizeof_atoms.restype = C.c_size_t
        _lib = L
    return _lib


def _p(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


def spline_build(dx, y):
    y = np.ascontiguousarray(y, dtype=np.float64)
    out = np.empty((len(y), 4))
    load().orc_spline_build(C.c_double(dx), _p(y), C.c_size_t(len(y)), _p(out))
    return out


def spline_eval(coeff, inv_dx, x):
    coeff = np.ascontiguousarray(coeff)
    L = load()
    return np.array([L.orc_spline_eval(_p(coeff), C.c_double(inv_dx), C.c_double(v)) for v in np.atleast_1d(x)])


def linear_eval(dx, y, x, reverse=False):
    y = np.ascontiguousarray(y, dtype=np.float64)
    L = load()
    f = L.orc_linear_reverse if reverse else L.orc_linear_eval
    return np.array([f(C.c_double(dx), _p(y), C.c_size_t(len(y)), C.c_double(v)) for v in np.atleast_1d(x)])


class Beta:
    def __init__(self, path=None, knots=None):
        L = load()
        if path is not None:
            self.h = L.orc_beta_load(str(path).encode())
        else:
            n_el, n_rho, dr, n_beta, drho, rc, rho_k, beta_k = knots
            rho_k = np.ascontiguousarray(rho_k, dtype=np.float64)
            beta_k = np.ascontiguousarray(beta_k, dtype=np.float64)
            self.h = L.orc_beta_from_knots(n_el, C.c_size_t(n_rho), C.c_double(dr), C.c_size_t(n_beta), C.c_double(drho),
                                           C.c_double(rc), _p(rho_k), _p(beta_k))
        if not self.h:
            raise RuntimeError("oracle: cannot load beta file %r" % (path,))
        dims = (C.c_longlong * 3)()
        scal = (C.c_double * 6)()
        L.orc_beta_info(C.c_void_p(self.h), dims, scal)
        self.n_elements, self.n_rho, self.n_beta = (int(d) for d in dims)
        (self.r_cutoff, self.r_cutoff_sq, self.rho_cutoff, self.inv_dr, self.inv_dr_sq, self.inv_drho) = (float(s) for s in scal)

    def table(self, kind, element=None):
        n = self.n_rho if kind < 2 else self.n_beta
        els = range(self.n_elements) if element is None else [element]
        out = np.empty((len(els), n, 4))
        for k, e in enumerate(els):
            ptr = load().orc_beta_table(C.c_void_p(self.h), kind, e)
            out[k] = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_double)), shape=(n, 4))
        return out if element is None else out[0]

    def eval(self, kind, e, x):
        L = load()
        f = {1: L.orc_beta_rho_r_sq, 2: L.orc_beta_alpha, 3: L.orc_beta_beta}[kind]
        return np.array([f(C.c_void_p(self.h), e, C.c_double(v)) for v in np.atleast_1d(x)])

    def __del__(self):
        try:
            load().orc_beta_free(C.c_void_p(self.h))
        except Exception:
            pass


class FDM:
    def __init__(self, nx=None, ny=None, nz=None, box=None, T_e=300.0, C_e=1.0, rho_e=1.0, kappa_e=1.0, path=None):
        L = load()
        if path is not None:
            self.h = L.orc_fdm_from_file(str(path).encode())
            if not self.h:
                raise RuntimeError("oracle: cannot load grid file %r" % (path,))
        else:
            b = np.ascontiguousarray(box, dtype=np.float64)
            self.h = L.orc_fdm_new(C.c_size_t(nx), C.c_size_t(ny), C.c_size_t(nz), _p(b), C.c_double(T_e), C.c_double(C_e),
                                   C.c_double(rho_e), C.c_double(kappa_e))
        # struct head: size_t nx, ny, nz, ntotal, steps
        head = C.cast(self.h, C.POINTER(C.c_size_t))
        self.nx, self.ny, self.nz, self.ntotal, self.steps = (int(head[i]) for i in range(5))

    def field(self, which):
        ptr = load().orc_fdm_field(C.c_void_p(self.h), which)
        return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_double)), shape=(self.ntotal,))

    def flags(self):
        L = load()
        fl = np.ctypeslib.as_array(C.cast(L.orc_fdm_flags(C.c_void_p(self.h)), C.POINTER(C.c_short)), shape=(self.ntotal,))
        td = np.ctypeslib.as_array(C.cast(L.orc_fdm_tdyn(C.c_void_p(self.h)), C.POINTER(C.c_ushort)), shape=(self.ntotal,))
        return fl, td

    def set_dt(self, dt): load().orc_fdm_set_dt(C.c_void_p(self.h), C.c_double(dt))
    def solve(self): load().orc_fdm_solve(C.c_void_p(self.h))
    def T_total(self): return load().orc_fdm_T_total(C.c_void_p(self.h))
    def index(self, x, y, z): return load().orc_fdm_index(C.c_void_p(self.h), C.c_double(x), C.c_double(y), C.c_double(z))

    def insert_energy(self, x, E):
        L = load()
        for p, e in zip(np.asarray(x), np.asarray(E)):
            L.orc_fdm_insert_energy(C.c_void_p(self.h), C.c_double(p[0]), C.c_double(p[1]), C.c_double(p[2]), C.c_double(e))

    def get_T(self, x):
        L = load()
        return np.array([L.orc_fdm_get_T(C.c_void_p(self.h), C.c_double(p[0]), C.c_double(p[1]), C.c_double(p[2])) for p in np.asarray(x)])

    def save_temperature(self, name, n): return load().orc_fdm_save_temperature(C.c_void_p(self.h), str(name).encode(), n)
    def save_state(self, path): return load().orc_fdm_save_state(C.c_void_p(self.h), str(path).encode())

    def last_substeps(self):
        # not exported separately: re-derive from the struct tail is brittle, so expose through a tiny helper
        return None

    def __del__(self):
        try:
            load().orc_fdm_free(C.c_void_p(self.h))
        except Exception:
            pass


class Fix:
    """The restated FixEPH hot path on one rank (periodic ghost images)."""

    def __init__(self, system, beta, fdm, flags, model=4, groupbit=1, type_map=None, dt=1e-4, boltz=8.617343e-5,
                 ftm2v=1.0 / 1.0364269e-4):
        L = load()
        self.sys, self.beta, self.fdm = system, beta, fdm
        ntypes = int(system.get("ntypes", 1))
        tm = np.ascontiguousarray(type_map if type_map is not None else list(range(ntypes)), dtype=np.int32)
        self.h = L.orc_fix_new(flags, model, groupbit, ntypes, _p(tm), C.c_double(dt), C.c_double(boltz), C.c_double(ftm2v),
                               C.c_void_p(beta.h), C.c_void_p(fdm.h))
        s = system
        self.nlocal, self.nghost = s["nlocal"], s["nghost"]
        self.x = np.ascontiguousarray(s["x"], dtype=np.float64).copy()
        self.v = np.ascontiguousarray(s["v"], dtype=np.float64).copy()
        self.f = np.ascontiguousarray(s["f"], dtype=np.float64).copy()
        self.type = np.ascontiguousarray(s["type"], dtype=np.int32)
        self.mask = np.ascontiguousarray(s["mask"], dtype=np.int32)
        self.owner = np.ascontiguousarray(s["ghost_owner"], dtype=np.int32)
        self.offsets = np.ascontiguousarray(s["offsets"], dtype=np.int64)
        self.neigh = np.ascontiguousarray(s["neigh"], dtype=np.int32)
        self.atoms = C.create_string_buffer(L.orc_sizeof_atoms())
        self._fill()

    def _fill(self):
        load().orc_atoms_fill(self.atoms, self.nlocal, self.nghost, _p(self.x), _p(self.v), _p(self.f), _p(self.type),
                              _p(self.mask), _p(self.owner), _p(self.offsets), _p(self.neigh))

    def set_dt(self, dt): load().orc_fix_set_dt(C.c_void_p(self.h), C.c_double(dt))

    def set_colour(self, tau0):
        """`fix eph/coloured/exp`: exponential memory kernel with time constant tau0 on both forces"""
        load().orc_fix_set_colour(C.c_void_p(self.h), C.c_double(tau0))

    def post_force(self, xi=None):
        xi = None if xi is None else np.ascontiguousarray(xi, dtype=np.float64)
        load().orc_post_force(C.c_void_p(self.h), self.atoms, _p(xi))

    def end_of_step(self): load().orc_end_of_step(C.c_void_p(self.h), self.atoms)

    def initial_integrate(self, mass):
        m = np.ascontiguousarray(np.concatenate([[0.0], np.atleast_1d(mass)]), dtype=np.float64)
        load().orc_initial_integrate(C.c_void_p(self.h), self.atoms, _p(m))

    def final_integrate(self, mass):
        m = np.ascontiguousarray(np.concatenate([[0.0], np.atleast_1d(mass)]), dtype=np.float64)
        load().orc_final_integrate(C.c_void_p(self.h), self.atoms, _p(m))

    def Ee(self): return load().orc_fix_Ee(C.c_void_p(self.h))

    def ptr(self, which):
        nl, nt = self.nlocal, self.nlocal + self.nghost
        shape = {0: (nt,), 1: (nt, 3), 2: (nt, 3), 3: (nt, 3), 4: (nt, 3), 5: (nt, 8), 6: (nt, 3), 7: (nt, 3)}[which]
        p = load().orc_fix_ptr(C.c_void_p(self.h), which)
        a = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_double)), shape=shape)
        return a if which == 0 else a[:nl]

    def __del__(self):
        try:
            load().orc_fix_free(C.c_void_p(self.h))
        except Exception:
            pass


class Kappa:
    """Restated EPH_kappa tables (eph_kappa.h)."""

    def __init__(self, path):
        L = load()
        self.h = L.orc_kappa_load(str(path).encode())
        if not self.h:
            raise RuntimeError("oracle: cannot load kappa file %r" % (path,))
        dims = (C.c_longlong * 4)()
        scal = (C.c_double * 5)()
        L.orc_kappa_info(C.c_void_p(self.h), dims, scal)
        self.n_elements, self.n_pairs, self.n_r, self.n_T = (int(d) for d in dims)
        self.r_cutoff, self.r_cutoff_sq, self.T_max, self.inv_dr_sq, self.dT = (float(v) for v in scal)

    def table(self, kind, e=0):
        """0 rho(r) [n_r][4], 1 rho(r^2) [n_r][4], 2 E(T) [n_T], 3 K(T) of slot e [n_T]"""
        shape = (self.n_r, 4) if kind < 2 else (self.n_T,)
        ptr = load().orc_kappa_table(C.c_void_p(self.h), kind, e)
        return np.array(np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_double)), shape=shape))

    def __del__(self):
        try:
            load().orc_kappa_free(C.c_void_p(self.h))
        except Exception:
            pass


class AtomicFix(Fix):
    """The restated FixEPHAtomic hot path on one rank (periodic ghost images), eph_oracle_atomic.c."""

    def __init__(self, system, beta, kappa, flags, groupbit=1, type_map_beta=None, type_map_kappa=None, dt=1e-4,
                 boltz=8.617343e-5, ftm2v=1.0 / 1.0364269e-4, inner_loops=0, T_init=300.0):
        L = load()
        self.sys, self.beta, self.kappa = system, beta, kappa
        ntypes = int(system.get("ntypes", 1))
        tb = np.ascontiguousarray(type_map_beta if type_map_beta is not None else list(range(ntypes)), dtype=np.int32)
        tk = np.ascontiguousarray(type_map_kappa if type_map_kappa is not None else list(range(ntypes)), dtype=np.int32)
        s = system
        self.nlocal, self.nghost = s["nlocal"], s["nghost"]
        self.x = np.ascontiguousarray(s["x"], dtype=np.float64).copy()
        self.v = np.ascontiguousarray(s["v"], dtype=np.float64).copy()
        self.f = np.ascontiguousarray(s["f"], dtype=np.float64).copy()
        self.type = np.ascontiguousarray(s["type"], dtype=np.int32)
        self.mask = np.ascontiguousarray(s["mask"], dtype=np.int32)
        self.owner = np.ascontiguousarray(s["ghost_owner"], dtype=np.int32)
        self.offsets = np.ascontiguousarray(s["offsets"], dtype=np.int64)
        self.neigh = np.ascontiguousarray(s["neigh"], dtype=np.int32)
        self.atoms = C.create_string_buffer(L.orc_sizeof_atoms())
        self._fill()
        self.h = L.orc_afix_new(flags, groupbit, ntypes, _p(tb), _p(tk), C.c_double(dt), C.c_double(boltz), C.c_double(ftm2v),
                                inner_loops, C.c_double(T_init), C.c_void_p(beta.h), C.c_void_p(kappa.h), self.atoms)

    def set_dt(self, dt): load().orc_afix_set_dt(C.c_void_p(self.h), C.c_double(dt))

    def post_force(self, xi=None):
        xi = None if xi is None else np.ascontiguousarray(xi, dtype=np.float64)
        load().orc_atomic_post_force(C.c_void_p(self.h), self.atoms, _p(xi))

    def end_of_step(self): load().orc_atomic_end_of_step(C.c_void_p(self.h), self.atoms)

    def initial_integrate(self, mass):
        m = np.ascontiguousarray(np.concatenate([[0.0], np.atleast_1d(mass)]), dtype=np.float64)
        load().orc_atomic_initial_integrate(C.c_void_p(self.h), self.atoms, _p(m))

    def final_integrate(self, mass):
        m = np.ascontiguousarray(np.concatenate([[0.0], np.atleast_1d(mass)]), dtype=np.float64)
        load().orc_atomic_final_integrate(C.c_void_p(self.h), self.atoms, _p(m))

    def Ee(self): return load().orc_afix_Ee(C.c_void_p(self.h))
    def Te(self): return load().orc_afix_Te(C.c_void_p(self.h))

    def ptr(self, which):
        """0 rho[nt] 1 w 2 xi 3 f_EPH 4 f_RNG [nl][3] 5 rho_a[nt] 6 E_a[nt] 7 dE_a[nl] 8 T_a[nl] 9 array[nl][12]"""
        nl, nt = self.nlocal, self.nlocal + self.nghost
        shape = {0: (nt,), 1: (nt, 3), 2: (nt, 3), 3: (nt, 3), 4: (nt, 3), 5: (nt,), 6: (nt, 2), 7: (nt,), 8: (nt,), 9: (nt, 12)}[which]
        p = load().orc_afix_ptr(C.c_void_p(self.h), which)
        a = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_double)), shape=shape)
        if which in (0, 5):
            return a
        if which == 6:
            return a[:, 0]
        return a[:nl]

    def set_energy(self, E):
        self.ptr(6)[: self.nlocal] = E

    def __del__(self):
        try:
            load().orc_afix_free(C.c_void_p(self.h))
        except Exception:
            pass


def xi_stream(seed, step, tags):
    tags = np.ascontiguousarray(tags, dtype=np.int64)
    out = np.empty((len(tags), 3))
    load().orc_xi_stream(C.c_uint64(seed), C.c_uint64(step), C.c_longlong(len(tags)), _p(tags), _p(out))
    return out
