"""TEST INFRASTRUCTURE: ctypes binding of the UNMODIFIED reference compiled into
oracle/_ref/libeph_ref.so (see oracle/ref/ref_driver.cpp).  The library is built
in the development container, where /root/reference exists, and travels to the
GPU box as a prebuilt file; nothing here reads /root/reference at run time."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PATH = os.path.join(HERE, "_ref", "libeph_ref.so")
_lib = None


def available():
    return os.path.exists(PATH)


def load():
    global _lib
    if _lib is None:
        L = C.CDLL(PATH)
        for n in ("ref_beta_load", "ref_fdm_new", "ref_fdm_from_file"):
            getattr(L, n).restype = C.c_void_p
        L.ref_fdm_T_total.restype = C.c_double
        _lib = L
    return _lib


def _p(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


def spline_build(dx, y):
    y = np.ascontiguousarray(y, dtype=np.float64)
    out = np.empty((len(y), 4))
    load().ref_spline_build(C.c_double(dx), _p(y), len(y), _p(out))
    return out


def spline_eval(dx, y, x):
    y = np.ascontiguousarray(y, dtype=np.float64)
    x = np.ascontiguousarray(x, dtype=np.float64)
    out = np.empty(len(x))
    load().ref_spline_eval(C.c_double(dx), _p(y), len(y), _p(x), len(x), _p(out))
    return out


def linear_eval(dx, y, x, reverse=False):
    y = np.ascontiguousarray(y, dtype=np.float64)
    x = np.ascontiguousarray(np.atleast_1d(x), dtype=np.float64)
    out = np.empty(len(x))
    load().ref_linear_eval(C.c_double(dx), _p(y), len(y), _p(x), len(x), _p(out), int(reverse))
    return out


def beta_tables(path):
    """The reference's EPH_Beta through the generic table reader of eph_b200.host"""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(HERE), "user-eph_b200"))
    from eph_b200.host import BetaTables
    return BetaTables(path=path, lib=load(), prefix="ref")


def beta_eval(tables, kind, e, x):
    x = np.ascontiguousarray(np.atleast_1d(x), dtype=np.float64)
    out = np.empty(len(x))
    load().ref_beta_eval(C.c_void_p(tables.h), kind, e, _p(x), len(x), _p(out))
    return out


class FDM:
    def __init__(self, nx=None, ny=None, nz=None, box=None, T_e=300.0, C_e=1.0, rho_e=1.0, kappa_e=1.0, path=None):
        L = load()
        if path is not None:
            self.h = L.ref_fdm_from_file(str(path).encode())
            if not self.h:
                raise RuntimeError("reference: cannot load grid file %r" % (path,))
        else:
            b = np.ascontiguousarray(box, dtype=np.float64)
            self.h = L.ref_fdm_new(nx, ny, nz, _p(b), C.c_double(T_e), C.c_double(C_e), C.c_double(rho_e), C.c_double(kappa_e))
        d = (C.c_longlong * 4)()
        L.ref_fdm_dims(C.c_void_p(self.h), d)
        self.nx, self.ny, self.nz, self.steps = (int(v) for v in d)
        self.ntotal = self.nx * self.ny * self.nz

    def get(self, which):
        out = np.empty(self.ntotal)
        load().ref_fdm_get(C.c_void_p(self.h), which, _p(out))
        return out

    def set(self, which, values):
        v = np.empty(self.ntotal)
        v[...] = values
        load().ref_fdm_set(C.c_void_p(self.h), which, _p(v))

    def get_flags(self):
        fl = np.empty(self.ntotal, dtype=np.int16)
        td = np.empty(self.ntotal, dtype=np.uint16)
        load().ref_fdm_get_flags(C.c_void_p(self.h), _p(fl), _p(td))
        return fl, td

    def set_flags(self, flag=None, tdyn=None):
        fl = None if flag is None else np.ascontiguousarray(flag, dtype=np.int16)
        td = None if tdyn is None else np.ascontiguousarray(tdyn, dtype=np.uint16)
        load().ref_fdm_set_flags(C.c_void_p(self.h), _p(fl), _p(td))

    def set_dt(self, dt): load().ref_fdm_set_dt(C.c_void_p(self.h), C.c_double(dt))
    def set_steps(self, s): load().ref_fdm_set_steps(C.c_void_p(self.h), C.c_longlong(s))
    def solve(self): load().ref_fdm_solve(C.c_void_p(self.h))
    def T_total(self): return load().ref_fdm_T_total(C.c_void_p(self.h))

    def insert_energy(self, x, E):
        x = np.ascontiguousarray(x, dtype=np.float64)
        E = np.ascontiguousarray(E, dtype=np.float64)
        load().ref_fdm_insert_energy(C.c_void_p(self.h), len(E), _p(x), _p(E))

    def get_T(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        out = np.empty(len(x))
        load().ref_fdm_get_T_at(C.c_void_p(self.h), len(x), _p(x), _p(out))
        return out

    def index(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        out = np.empty(len(x), dtype=np.int64)
        load().ref_fdm_index_at(C.c_void_p(self.h), len(x), _p(x), _p(out))
        return out

    def save_temperature(self, name, n): load().ref_fdm_save_temperature(C.c_void_p(self.h), str(name).encode(), n)
    def save_state(self, path): load().ref_fdm_save_state(C.c_void_p(self.h), str(path).encode())

    def __del__(self):
        try:
            load().ref_fdm_free(C.c_void_p(self.h))
        except Exception:
            pass


def fix_driver(system, fix_args, dt=1e-4, mass=None):
    """The unmodified FixEPH inside the LAMMPS stand-in."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(HERE), "user-eph_b200"))
    from eph_b200.host import FixDriver
    return FixDriver(system, fix_args, dt=dt, lib=load(), prefix="ref", mass=mass)


# ---------------------------------------------------------------------------
# the unmodified `fix eph/atomic` (oracle/_ref/libeph_atomic_ref.so, oracle/ref/ref_atomic_driver.cpp)
# ---------------------------------------------------------------------------
PATH_ATOMIC = os.path.join(HERE, "_ref", "libeph_atomic_ref.so")
_alib = None


def atomic_available():
    return os.path.exists(PATH_ATOMIC)


def load_atomic():
    global _alib
    if _alib is None:
        L = C.CDLL(PATH_ATOMIC)
        L.refa_kappa_load.restype = C.c_void_p
        _alib = L
    return _alib


def atomic_fix_driver(system, fix_args, dt=1e-4, mass=None):
    """The unmodified FixEPHAtomic inside the LAMMPS stand-in."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(HERE), "user-eph_b200"))
    from eph_b200.host import FixDriver
    return FixDriver(system, fix_args, dt=dt, lib=load_atomic(), prefix="refa", mass=mass)


class Kappa:
    """The reference's EPH_kappa tables (eph_kappa.h) at full precision."""

    def __init__(self, path):
        L = load_atomic()
        self.h = L.refa_kappa_load(str(path).encode())
        if not self.h:
            raise RuntimeError("reference: cannot load kappa file %r" % (path,))
        dims = (C.c_longlong * 4)()
        scal = (C.c_double * 5)()
        L.refa_kappa_info(C.c_void_p(self.h), dims, scal)
        self.n_elements, self.n_pairs, self.n_r, self.n_T = (int(d) for d in dims)
        self.r_cutoff, self.r_cutoff_sq, self.T_max, self.inv_dr_sq, self.dT = (float(v) for v in scal)

    def table(self, kind, e=0):
        """0 rho(r) [n_r][4], 1 rho(r^2) [n_r][4], 2 E(T) [n_T], 3 K(T) of slot e [n_T]"""
        out = np.empty((self.n_r, 4)) if kind < 2 else np.empty(self.n_T)
        load_atomic().refa_kappa_table(C.c_void_p(self.h), kind, e, _p(out))
        return out

    def __del__(self):
        try:
            load_atomic().refa_kappa_free(C.c_void_p(self.h))
        except Exception:
            pass


# ---------------------------------------------------------------------------
# the unmodified `fix eph/coloured/exp` (oracle/_ref/libeph_coloured_ref.so, oracle/ref/ref_coloured_driver.cpp)
# ---------------------------------------------------------------------------
PATH_COLOURED = os.path.join(HERE, "_ref", "libeph_coloured_ref.so")
_clib = None


def coloured_available():
    return os.path.exists(PATH_COLOURED)


def coloured_fix_driver(system, fix_args, dt=1e-4, mass=None):
    """The unmodified FixEPHColouredExp inside the LAMMPS stand-in; probes 5 / 6 are f_dis / f_sto [nlocal][3]."""
    global _clib
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(HERE), "user-eph_b200"))
    from eph_b200.host import FixDriver
    if _clib is None:
        _clib = C.CDLL(PATH_COLOURED)
    return FixDriver(system, fix_args, dt=dt, lib=_clib, prefix="refc", mass=mass)
