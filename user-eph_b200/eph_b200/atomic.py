"""ctypes binding of the `fix eph/atomic` path: the C ABI in include/eph_b200_atomic.h (libeph_b200.so) and the host
side of FixEPHAtomicB200 (libeph_b200_atomic_fix.so: `.kappa` tables, the shim-driven fix).

Array arguments may be numpy arrays (host memspace) or torch CUDA tensors (device memspace), not mixed in one call."""
import ctypes as C

import numpy as np

from . import lib as _lib
from ._paths import lib_path
from .lib import DEVICE, HOST, EphError, _ptr, _space, c_double_p, c_int_p  # noqa: F401


class AtomicConfig(C.Structure):
    _fields_ = [("device", C.c_int), ("ntypes", C.c_int), ("type_map_beta", c_int_p), ("type_map_kappa", c_int_p),
                ("groupbit", C.c_int), ("flags", C.c_int), ("seed", C.c_ulonglong), ("inner_loops", C.c_int),
                ("stream", C.c_void_p)]


# every symbol include/eph_b200_atomic.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "eph_b200_atomic_create": (C.c_int, [C.POINTER(AtomicConfig), C.POINTER(C.c_void_p)]),
    "eph_b200_atomic_destroy": (C.c_int, [C.c_void_p]),
    "eph_b200_atomic_last_error": (C.c_char_p, [C.c_void_p]),
    "eph_b200_atomic_create_error": (C.c_char_p, []),
    "eph_b200_atomic_launch_count": (C.c_longlong, [C.c_void_p]),
    "eph_b200_atomic_synchronize": (C.c_int, [C.c_void_p]),
    "eph_b200_atomic_set_beta_tables": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_void_p, C.c_int, C.c_double,
                                                  C.c_void_p, C.c_void_p, C.c_double, C.c_double]),
    "eph_b200_atomic_set_kappa_tables": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_void_p, C.c_double,
                                                   C.c_int, C.c_double, C.c_void_p, C.c_void_p]),
    "eph_b200_atomic_set_dt": (C.c_int, [C.c_void_p, C.c_double, C.c_double]),
    "eph_b200_atomic_set_atoms": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]),
    "eph_b200_atomic_set_neighbors_csr": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int]),
    "eph_b200_atomic_init_energy": (C.c_int, [C.c_void_p, C.c_double]),
    "eph_b200_atomic_set_energy": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    "eph_b200_atomic_get_energy": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    "eph_b200_atomic_post_force": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.c_int]),
    "eph_b200_atomic_set_comm_mode": (C.c_int, [C.c_void_p, C.c_int]),
    "eph_b200_atomic_post_force_begin": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.c_int]),
    "eph_b200_atomic_post_force_mid": (C.c_int, [C.c_void_p]),
    "eph_b200_atomic_post_force_end": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    "eph_b200_atomic_heat_loops": (C.c_int, [C.c_void_p]),
    "eph_b200_atomic_heat_begin": (C.c_int, [C.c_void_p]),
    "eph_b200_atomic_heat_end": (C.c_int, [C.c_void_p]),
    "eph_b200_atomic_pack_forward": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "eph_b200_atomic_unpack_forward": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "eph_b200_atomic_end_of_step": (C.c_int, [C.c_void_p, c_double_p, c_double_p]),
    "eph_b200_atomic_summary": (C.c_int, [C.c_void_p, c_double_p, c_double_p]),
    "eph_b200_atomic_get_peratom": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    "eph_b200_atomic_get_probe": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
}

_declared = False
_fix = None


def declare(L):
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    return L


def load():
    """libeph_b200.so with the atomic prototypes declared.  Raises if the library is missing: no fallback."""
    global _declared
    L = _lib.load()
    if not _declared:
        declare(L)
        _declared = True
    return L


def load_fix_lib():
    global _fix
    if _fix is None:
        load()  # libeph_b200_atomic_fix.so links libeph_b200.so
        _fix = C.CDLL(lib_path("atomic_fix"))
        _fix.ephk_last_error.restype = C.c_char_p
        _fix.ephk_kappa_load.restype = C.c_void_p
        _fix.ephk_linear.restype = C.c_double
    return _fix


class KappaTables:
    """`.kappa` tables built by the product's host code (fix/eph_kappa_tables.h)."""

    def __init__(self, path):
        self.lib = load_fix_lib()
        self.h = self.lib.ephk_kappa_load(str(path).encode())
        if not self.h:
            raise RuntimeError(self.lib.ephk_last_error().decode())
        dims = (C.c_longlong * 4)()
        scal = (C.c_double * 5)()
        self.lib.ephk_kappa_info(C.c_void_p(self.h), dims, scal)
        self.n_elements, self.n_pairs, self.n_r, self.n_T = (int(d) for d in dims)
        self.r_cutoff, self.r_cutoff_sq, self.T_max, self.inv_dr_sq, self.dT = (float(v) for v in scal)

    def name(self, e):
        buf = C.create_string_buffer(64)
        self.lib.ephk_kappa_name(C.c_void_p(self.h), e, buf, 64)
        return buf.value.decode()

    def table(self, kind, e=0):
        """0 rho(r) [n_r][4], 1 rho(r^2) [n_r][4], 2 E(T) [n_T], 3 K(T) of slot e [n_T]"""
        out = np.empty((self.n_r, 4)) if kind < 2 else np.empty(self.n_T)
        self.lib.ephk_kappa_table(C.c_void_p(self.h), kind, e, C.c_void_p(out.ctypes.data))
        return out

    def linear(self, e, x, reverse=False):
        return np.array([self.lib.ephk_linear(C.c_void_p(self.h), e, C.c_double(v), int(reverse)) for v in np.atleast_1d(x)])

    def __del__(self):
        try:
            if self.h:
                self.lib.ephk_kappa_free(C.c_void_p(self.h))
        except Exception:
            pass


def fix_driver(system, fix_args, dt=1e-4, mass=None, lib=None):
    """FixEPHAtomicB200 inside the LAMMPS stand-in (same calls as oracle.reference.atomic_fix_driver)."""
    from .host import FixDriver
    return FixDriver(system, fix_args, dt=dt, lib=lib or load_fix_lib(), prefix="b200a", mass=mass)


class AtomicEngine:
    """One eph_b200_atomic_handle: the device engine behind FixEPHAtomicB200."""

    def __init__(self, type_map_beta, type_map_kappa, flags, groupbit=1, seed=12345, inner_loops=0, device=0, stream=None,
                 lib=None):
        self.lib = lib or load()   # `lib`: tests/test_atomic_emulated.py passes the host build of the same source
        tb = np.ascontiguousarray(type_map_beta, dtype=np.int32)
        tk = np.ascontiguousarray(type_map_kappa, dtype=np.int32)
        assert len(tb) == len(tk)
        cfg = AtomicConfig(device, len(tb), tb.ctypes.data_as(c_int_p), tk.ctypes.data_as(c_int_p), groupbit, flags, seed,
                           inner_loops, C.c_void_p(stream) if stream else None)
        h = C.c_void_p()
        rc = self.lib.eph_b200_atomic_create(C.byref(cfg), C.byref(h))
        if rc != 0:
            raise EphError("eph_b200_atomic_create failed (%d): %s" % (rc, self.lib.eph_b200_atomic_create_error().decode()))
        self.h = h
        self.flags = flags
        self.nlocal = self.nghost = 0

    def _check(self, rc):
        if rc != 0:
            raise EphError("eph_b200_atomic error %d: %s" % (rc, self.lib.eph_b200_atomic_last_error(self.h).decode()))

    def close(self):
        if getattr(self, "h", None):
            self.lib.eph_b200_atomic_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_tables_from(self, beta, kappa):
        """beta: eph_b200.host.BetaTables; kappa: KappaTables (or any object with the same table()/attributes)"""
        t1, t2, t3 = (np.ascontiguousarray(beta.table(k), dtype=np.float64) for k in (1, 2, 3))
        self._check(self.lib.eph_b200_atomic_set_beta_tables(self.h, beta.n_elements, beta.n_rho, beta.inv_dr_sq, t1.ctypes.data,
                                                             beta.n_beta, beta.inv_drho, t2.ctypes.data, t3.ctypes.data,
                                                             beta.r_cutoff_sq, beta.rho_cutoff))
        kr = np.ascontiguousarray(np.stack([kappa.table(1, e) for e in range(kappa.n_elements)]), dtype=np.float64)
        kE = np.ascontiguousarray(np.stack([kappa.table(2, e) for e in range(kappa.n_elements)]), dtype=np.float64)
        kK = np.ascontiguousarray(np.stack([kappa.table(3, p) for p in range(kappa.n_pairs)]), dtype=np.float64)
        self._check(self.lib.eph_b200_atomic_set_kappa_tables(self.h, kappa.n_elements, kappa.n_pairs, kappa.n_r, kappa.inv_dr_sq,
                                                              kr.ctypes.data, kappa.r_cutoff_sq, kappa.n_T, kappa.dT,
                                                              kE.ctypes.data, kK.ctypes.data))

    def set_dt(self, dt, boltz=8.617343e-5):
        self._check(self.lib.eph_b200_atomic_set_dt(self.h, dt, boltz))

    def set_atoms(self, nlocal, nghost, type, mask, tag, ghost_owner):
        ps = [_ptr(type), _ptr(mask), _ptr(tag), _ptr(ghost_owner)]
        self._check(self.lib.eph_b200_atomic_set_atoms(self.h, nlocal, nghost, ps[0][0], ps[1][0], ps[2][0], ps[3][0], _space(*ps)))
        self.nlocal, self.nghost = nlocal, nghost
        self._keep_atoms = (type, mask, tag)   # device memspace aliases the caller's buffers

    def set_neighbors(self, offsets, neigh):
        ps = [_ptr(offsets), _ptr(neigh)]
        self._check(self.lib.eph_b200_atomic_set_neighbors_csr(self.h, self.nlocal, ps[0][0], ps[1][0], _space(*ps)))
        self._keep_list = (offsets, neigh)

    def init_energy(self, T_init):
        self._check(self.lib.eph_b200_atomic_init_energy(self.h, T_init))

    def set_energy(self, E):
        p = _ptr(E)
        self._check(self.lib.eph_b200_atomic_set_energy(self.h, p[0], p[1]))

    def get_energy(self):
        out = np.empty(self.nlocal)
        self._check(self.lib.eph_b200_atomic_get_energy(self.h, out.ctypes.data, HOST))
        return out

    def post_force(self, x, v, f, xi=None, step=0):
        ps = [_ptr(x), _ptr(v), _ptr(f), _ptr(xi)]
        self._check(self.lib.eph_b200_atomic_post_force(self.h, ps[0][0], ps[1][0], ps[2][0], ps[3][0], step, _space(*ps)))

    def end_of_step(self):
        e, t = C.c_double(), C.c_double()
        self._check(self.lib.eph_b200_atomic_end_of_step(self.h, C.byref(e), C.byref(t)))
        return e.value, t.value

    def summary(self):
        e, t = C.c_double(), C.c_double()
        self._check(self.lib.eph_b200_atomic_summary(self.h, C.byref(e), C.byref(t)))
        return e.value, t.value

    def peratom(self):
        out = np.empty((self.nlocal, 12))
        self._check(self.lib.eph_b200_atomic_get_peratom(self.h, out.ctypes.data, HOST))
        return out

    def probe(self, which):
        nt = self.nlocal + self.nghost
        n = nt if which in (0, 5, 6) else self.nlocal if which in (7, 8) else 3 * self.nlocal
        out = np.empty(n)
        self._check(self.lib.eph_b200_atomic_get_probe(self.h, which, out.ctypes.data))
        return out.reshape(-1, 3) if which in (1, 2, 3, 4) else out

    def synchronize(self):
        self._check(self.lib.eph_b200_atomic_synchronize(self.h))

    def launch_count(self):
        return self.lib.eph_b200_atomic_launch_count(self.h)
