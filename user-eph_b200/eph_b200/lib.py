"""ctypes binding of the C ABI in include/eph_b200.h (libeph_b200.so).

Array arguments may be numpy arrays (host memspace) or torch CUDA tensors
(device memspace); the two must not be mixed inside one call.
"""
import ctypes as C

import numpy as np

from ._paths import lib_path

HOST, DEVICE = 0, 1
FRICTION, RANDOM, FDM, NOINT, NOFRICTION, NORANDOM = 1, 2, 4, 8, 16, 32

c_double_p = C.POINTER(C.c_double)
c_int_p = C.POINTER(C.c_int)
c_i64_p = C.POINTER(C.c_int64)


class Config(C.Structure):
    _fields_ = [("device", C.c_int), ("ntypes", C.c_int), ("type_map", c_int_p), ("groupbit", C.c_int),
                ("flags", C.c_int), ("model", C.c_int), ("seed", C.c_ulonglong), ("rank", C.c_int),
                ("nranks", C.c_int), ("stream", C.c_void_p)]


# every symbol include/eph_b200.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "eph_b200_version": (C.c_int, []),
    "eph_b200_device_count": (C.c_int, [c_int_p]),
    "eph_b200_create": (C.c_int, [C.POINTER(Config), C.POINTER(C.c_void_p)]),
    "eph_b200_destroy": (C.c_int, [C.c_void_p]),
    "eph_b200_last_error": (C.c_char_p, [C.c_void_p]),
    "eph_b200_create_error": (C.c_char_p, []),
    "eph_b200_set_tables": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_void_p, C.c_int, C.c_double,
                                      C.c_void_p, C.c_void_p, C.c_double, C.c_double]),
    "eph_b200_set_rho_r_table": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_void_p]),
    "eph_b200_set_grid": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "eph_b200_set_grid_tables": (C.c_int, [C.c_void_p, C.c_int, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]),
    "eph_b200_get_grid": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "eph_b200_put_grid": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "eph_b200_mean_T": (C.c_int, [C.c_void_p, c_double_p]),
    "eph_b200_last_substeps": (C.c_int, [C.c_void_p, c_int_p]),
    "eph_b200_set_dt": (C.c_int, [C.c_void_p, C.c_double, C.c_double]),
    "eph_b200_set_colour": (C.c_int, [C.c_void_p, C.c_double]),
    "eph_b200_get_colour_state": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]),
    "eph_b200_set_colour_state": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]),
    "eph_b200_set_precision": (C.c_int, [C.c_void_p, C.c_int]),
    "eph_b200_get_precision": (C.c_int, [C.c_void_p, c_int_p, c_double_p, c_double_p]),
    "eph_b200_set_skin": (C.c_int, [C.c_void_p, C.c_double, C.c_double]),
    "eph_b200_list_stats": (C.c_int, [C.c_void_p, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong)]),
    "eph_b200_set_atoms": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]),
    "eph_b200_set_neighbors_csr": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int]),
    "eph_b200_set_neighbors_lammps": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "eph_b200_build_neighbors": (C.c_int, [C.c_void_p, C.c_void_p, C.c_double, C.c_int]),
    "eph_b200_get_neighbors": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_longlong)]),
    "eph_b200_post_force": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.c_int]),
    "eph_b200_end_of_step": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, c_double_p, C.c_int]),
    "eph_b200_post_force_begin": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.c_int]),
    "eph_b200_pack_ghost_payload": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "eph_b200_unpack_ghost_payload": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "eph_b200_post_force_end": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    "eph_b200_end_of_step_begin": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]),
    "eph_b200_end_of_step_end": (C.c_int, [C.c_void_p, c_double_p]),
    "eph_b200_bind_grid_source": (C.c_int, [C.c_void_p, C.c_void_p]),
    "eph_b200_grid_plan_substeps": (C.c_int, [C.c_void_p, c_int_p]),
    "eph_b200_grid_substep": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "eph_b200_end_of_step_end_external": (C.c_int, [C.c_void_p, c_double_p]),
    "eph_b200_grid_device_ptr": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_void_p)]),
    "eph_b200_comm_get_id": (C.c_int, [C.c_void_p]),
    "eph_b200_comm_init": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int]),
    "eph_b200_comm_transport": (C.c_int, [C.c_void_p]),
    "eph_b200_set_ghost_map": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "eph_b200_exchange_ghosts": (C.c_int, [C.c_void_p]),
    "eph_b200_set_grid_sharding": (C.c_int, [C.c_void_p, C.c_int]),
    "eph_b200_reduce_and_solve": (C.c_int, [C.c_void_p, c_double_p]),
    "eph_b200_set_grid_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "eph_b200_set_comm_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "eph_b200_set_boundary_atoms": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int]),
    "eph_b200_initial_integrate": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double,
                                             C.c_double, C.c_int]),
    "eph_b200_final_integrate": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_int]),
    "eph_b200_refresh_ghosts": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "eph_b200_resident_upload": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "eph_b200_resident_initial_integrate": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_void_p, C.c_longlong]),
    "eph_b200_resident_post_force": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong]),
    "eph_b200_resident_final_integrate": (C.c_int, [C.c_void_p, C.c_void_p, C.c_double, C.c_void_p]),
    "eph_b200_resident_end_of_step": (C.c_int, [C.c_void_p, c_double_p]),
    "eph_b200_resident_get": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "eph_b200_get_peratom": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    "eph_b200_get_probe": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "eph_b200_pack_forward": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "eph_b200_unpack_forward": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "eph_b200_set_profiling": (C.c_int, [C.c_void_p, C.c_int]),
    "eph_b200_kernel_times": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_char_p), c_double_p, C.POINTER(C.c_longlong)]),
    "eph_b200_synchronize": (C.c_int, [C.c_void_p]),
    "eph_b200_launch_count": (C.c_longlong, [C.c_void_p]),
    "eph_b200_status_word": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint)]),
}

_lib = None


def load():
    """Load libeph_b200.so and declare every prototype.  Raises if the library is missing: no fallback."""
    global _lib
    if _lib is None:
        lib = C.CDLL(lib_path("engine"), mode=C.RTLD_GLOBAL)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


class EphError(RuntimeError):
    pass


def _ptr(a):
    """(address, memspace) of a numpy array or torch tensor; None -> (None, HOST)."""
    if a is None:
        return None, None
    if isinstance(a, np.ndarray):
        if not a.flags["C_CONTIGUOUS"]:
            raise ValueError("array must be C-contiguous")
        return a.ctypes.data, HOST
    # torch tensor
    if not a.is_contiguous():
        raise ValueError("tensor must be contiguous")
    return a.data_ptr(), (DEVICE if a.is_cuda else HOST)


def _space(*arrs):
    spaces = {s for _, s in arrs if s is not None}
    if len(spaces) > 1:
        raise ValueError("cannot mix host and device arrays in one call")
    return spaces.pop() if spaces else HOST


class _DeviceArray:
    """fp64 device memory owned by the engine, exposed through the CUDA array interface (zero-copy torch views)"""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": "<f8", "data": (int(ptr), False), "version": 2,
                                         "strides": None}


class Engine:
    """One eph_b200_handle: the device engine behind FixEPHB200."""

    def __init__(self, type_map, flags, model=4, groupbit=1, seed=12345, device=0, rank=0, nranks=1, stream=None):
        self.lib = load()
        tm = np.ascontiguousarray(type_map, dtype=np.int32)
        cfg = Config(device, len(tm), tm.ctypes.data_as(c_int_p), groupbit, flags, model, seed, rank, nranks,
                     C.c_void_p(stream) if stream else None)
        h = C.c_void_p()
        rc = self.lib.eph_b200_create(C.byref(cfg), C.byref(h))
        if rc != 0:
            raise EphError("eph_b200_create failed (%d): %s" % (rc, self.lib.eph_b200_create_error().decode()))
        self.h = h
        self.flags = flags
        self.model = model
        self.nlocal = self.nghost = 0
        self.ncell = 0

    def _check(self, rc):
        if rc != 0:
            raise EphError("eph_b200 error %d: %s" % (rc, self.lib.eph_b200_last_error(self.h).decode()))

    def close(self):
        if getattr(self, "h", None):
            self.lib.eph_b200_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- set-up ---------------------------------------------------------------
    def set_tables(self, n_elements, n_rho, inv_dr_sq, rho_r_sq, n_beta, inv_drho, alpha, beta, r_cutoff_sq, rho_cutoff):
        rho_r_sq, alpha, beta = (np.ascontiguousarray(t, dtype=np.float64) for t in (rho_r_sq, alpha, beta))
        self._check(self.lib.eph_b200_set_tables(self.h, n_elements, n_rho, inv_dr_sq, rho_r_sq.ctypes.data, n_beta,
                                                 inv_drho, alpha.ctypes.data, beta.ctypes.data, r_cutoff_sq, rho_cutoff))

    def set_tables_from(self, tables):
        """tables: eph_b200.host.BetaTables"""
        self.set_tables(tables.n_elements, tables.n_rho, tables.inv_dr_sq, tables.table(1), tables.n_beta,
                        tables.inv_drho, tables.table(2), tables.table(3), tables.r_cutoff_sq, tables.rho_cutoff)
        if self.model == 2:   # PRB evaluates rho(r) per pair (fix_eph.cpp:530)
            t = np.ascontiguousarray(tables.table(0), dtype=np.float64)
            self._check(self.lib.eph_b200_set_rho_r_table(self.h, tables.n_elements, tables.n_rho, C.c_double(tables.inv_dr),
                                                          t.ctypes.data))

    def set_grid(self, nx, ny, nz, box, T_e, rho_e, C_e, kappa_e, S_e=None, flag=None, t_dyn=None, steps=1):
        n = nx * ny * nz

        def field(v):
            a = np.empty(n, dtype=np.float64)
            a[...] = v
            return a

        box = np.ascontiguousarray(box, dtype=np.float64)
        T_e, rho_e, C_e, kappa_e = field(T_e), field(rho_e), field(C_e), field(kappa_e)
        S_e = field(0.0 if S_e is None else S_e)
        fl = None if flag is None else np.ascontiguousarray(flag, dtype=np.int16)
        td = None if t_dyn is None else np.ascontiguousarray(t_dyn, dtype=np.uint16)
        self._check(self.lib.eph_b200_set_grid(self.h, nx, ny, nz, box.ctypes.data, steps, T_e.ctypes.data, S_e.ctypes.data,
                                               rho_e.ctypes.data, C_e.ctypes.data, kappa_e.ctypes.data,
                                               None if fl is None else fl.ctypes.data, None if td is None else td.ctypes.data))
        self.ncell = n
        self.grid_shape = (nx, ny, nz)

    def set_grid_tables(self, dT, C_e_T, kappa_e_T, E_e_T):
        C_e_T, kappa_e_T, E_e_T = (np.ascontiguousarray(t, dtype=np.float64) for t in (C_e_T, kappa_e_T, E_e_T))
        self._check(self.lib.eph_b200_set_grid_tables(self.h, len(E_e_T), dT, C_e_T.ctypes.data, kappa_e_T.ctypes.data,
                                                      E_e_T.ctypes.data))

    def get_grid(self, which=0):
        out = np.empty(self.ncell, dtype=np.float64)
        self._check(self.lib.eph_b200_get_grid(self.h, which, out.ctypes.data))
        return out

    def put_grid(self, which, values):
        v = np.ascontiguousarray(values, dtype=np.float64)
        assert v.size == self.ncell
        self._check(self.lib.eph_b200_put_grid(self.h, which, v.ctypes.data))

    def mean_T(self):
        out = C.c_double()
        self._check(self.lib.eph_b200_mean_T(self.h, C.byref(out)))
        return out.value

    def last_substeps(self):
        out = C.c_int()
        self._check(self.lib.eph_b200_last_substeps(self.h, C.byref(out)))
        return out.value

    def set_dt(self, dt, boltz=8.617343e-5):
        self._check(self.lib.eph_b200_set_dt(self.h, dt, boltz))

    def set_colour(self, tau0):
        """`fix eph/coloured/exp`: exponential memory kernel with time constant tau0 on both forces (0: off)"""
        self._check(self.lib.eph_b200_set_colour(self.h, tau0))

    def colour_state(self):
        fd, fs = np.empty((self.nlocal, 3)), np.empty((self.nlocal, 3))
        self._check(self.lib.eph_b200_get_colour_state(self.h, fd.ctypes.data, fs.ctypes.data, HOST))
        return fd, fs

    def set_colour_state(self, f_dis, f_sto):
        ps = [_ptr(f_dis), _ptr(f_sto)]
        self._check(self.lib.eph_b200_set_colour_state(self.h, ps[0][0], ps[1][0], _space(*ps)))

    def set_precision(self, packed=True):
        """gather records of the sweeps: packed (16-byte positions / vectors, default) or fp64"""
        self._check(self.lib.eph_b200_set_precision(self.h, int(bool(packed))))

    def precision(self):
        on, period, quantum = C.c_int(), C.c_double(), C.c_double()
        self._check(self.lib.eph_b200_get_precision(self.h, C.byref(on), C.byref(period), C.byref(quantum)))
        return {"packed": bool(on.value), "period": period.value, "position_quantum": quantum.value}

    def set_skin(self, skin, inner_skin=-1.0):
        self._check(self.lib.eph_b200_set_skin(self.h, skin, inner_skin))

    def list_stats(self):
        a, b = C.c_longlong(), C.c_longlong()
        self._check(self.lib.eph_b200_list_stats(self.h, C.byref(a), C.byref(b)))
        return {"inner_builds": a.value, "fallback_steps": b.value}

    def set_atoms(self, nlocal, nghost, type, mask, tag=None, ghost_owner=None):
        ps = [_ptr(type), _ptr(mask), _ptr(tag), _ptr(ghost_owner)]
        sp = _space(*ps)
        self._check(self.lib.eph_b200_set_atoms(self.h, nlocal, nghost, ps[0][0], ps[1][0], ps[2][0], ps[3][0], sp))
        self.nlocal, self.nghost = nlocal, nghost

    def set_neighbors(self, offsets, neigh):
        ps = [_ptr(offsets), _ptr(neigh)]
        self._check(self.lib.eph_b200_set_neighbors_csr(self.h, self.nlocal, ps[0][0], ps[1][0], _space(*ps)))
        self._keep = (offsets, neigh)  # device memspace aliases the caller's buffers

    def build_neighbors(self, x, cutoff):
        """full list built on the device from positions (instead of uploading LAMMPS' list); x = None: the positions the
        engine keeps itself (resident mode)"""
        p = _ptr(x) if x is not None else (None, HOST)
        self._check(self.lib.eph_b200_build_neighbors(self.h, p[0], cutoff, p[1]))

    def get_neighbors_count(self):
        """entries of the full list in use"""
        n = C.c_longlong()
        self._check(self.lib.eph_b200_get_neighbors(self.h, None, None, C.byref(n)))
        return n.value

    def get_neighbors(self):
        n = C.c_longlong()
        self._check(self.lib.eph_b200_get_neighbors(self.h, None, None, C.byref(n)))
        off = np.empty(self.nlocal + 1, dtype=np.int64)
        ne = np.empty(max(n.value, 1), dtype=np.int32)
        self._check(self.lib.eph_b200_get_neighbors(self.h, off.ctypes.data, ne.ctypes.data, None))
        return off, ne[: n.value]

    # -- per step -------------------------------------------------------------
    def post_force(self, x, v, f, xi=None, step=0):
        ps = [_ptr(x), _ptr(v), _ptr(f), _ptr(xi)]
        self._check(self.lib.eph_b200_post_force(self.h, ps[0][0], ps[1][0], ps[2][0], ps[3][0], step, _space(*ps)))

    # multi-rank halves (device tensors)
    def post_force_begin(self, x, v, xi=None, step=0):
        ps = [_ptr(x), _ptr(v), _ptr(xi)]
        self._check(self.lib.eph_b200_post_force_begin(self.h, ps[0][0], ps[1][0], ps[2][0], step, _space(*ps)))

    def pack_ghost_payload(self, index, buf):
        self._check(self.lib.eph_b200_pack_ghost_payload(self.h, index.numel(), index.data_ptr(), buf.data_ptr()))

    def unpack_ghost_payload(self, index, buf):
        self._check(self.lib.eph_b200_unpack_ghost_payload(self.h, index.numel(), index.data_ptr(), buf.data_ptr()))

    def post_force_end(self, f):
        p = _ptr(f)
        self._check(self.lib.eph_b200_post_force_end(self.h, p[0], p[1]))

    def end_of_step_begin(self, x, v):
        ps = [_ptr(x), _ptr(v)]
        self._check(self.lib.eph_b200_end_of_step_begin(self.h, ps[0][0], ps[1][0], _space(*ps)))

    def end_of_step_end(self, want_energy=True, external=False):
        """external: the caller has run the grid solve itself (sharded solve), only the bookkeeping is left"""
        e = C.c_double()
        fn = self.lib.eph_b200_end_of_step_end_external if external else self.lib.eph_b200_end_of_step_end
        self._check(fn(self.h, C.byref(e) if want_energy else None))
        return e.value if want_energy else None

    # sharded grid solve (eph_harness.parallel.sharded_grid_solve drives these between end_of_step_begin and
    # end_of_step_end(external=True))
    def grid_plan_substeps(self):
        n = C.c_int()
        self._check(self.lib.eph_b200_grid_plan_substeps(self.h, C.byref(n)))
        return n.value

    def grid_substep(self, z_begin, z_end):
        self._check(self.lib.eph_b200_grid_substep(self.h, z_begin, z_end))

    def grid_tensor(self, which=0):
        """torch view (no copy) of a grid field in device memory; for T_e (which = 0) it names the CURRENT buffer of the
        double-buffered solve, so ask again after every sub-step"""
        import torch
        p = C.c_void_p()
        self._check(self.lib.eph_b200_grid_device_ptr(self.h, which, C.byref(p)))
        views = self.__dict__.setdefault("_grid_views", {})
        if p.value not in views:
            views[p.value] = torch.as_tensor(_DeviceArray(p.value, self.ncell), device=torch.device("cuda", torch.cuda.current_device()))
        return views[p.value]

    # -- multi-rank data plane inside the engine (NCCL) ------------------------
    @staticmethod
    def comm_get_id():
        """128 bytes (ncclUniqueId) created on the calling rank, for the caller to broadcast"""
        buf = (C.c_char * 128)()
        rc = load().eph_b200_comm_get_id(buf)
        if rc != 0:
            raise EphError("eph_b200_comm_get_id failed (%d): %s" % (rc, load().eph_b200_create_error().decode()))
        return bytes(buf)

    def comm_init(self, id_bytes, rank, nranks):
        buf = (C.c_char * 128).from_buffer_copy(id_bytes)
        self._check(self.lib.eph_b200_comm_init(self.h, buf, rank, nranks))

    def comm_transport(self):
        """0 no communicator, 1 NCCL send / receive, 2 NVLink peer memory (eph_p2p.cuh)"""
        return int(self.lib.eph_b200_comm_transport(self.h))

    def set_ghost_map(self, plan):
        """plan: eph_harness.parallel.ExchangePlan (who holds which of my atoms as ghosts, who fills which of my ghost slots)"""
        cached = getattr(plan, "_ghost_map_arrays", None)   # a plan does not change: flatten it once
        if cached is None:
            peers = [r for r in range(plan.world) if r != plan.rank and (plan.send_counts[r] or plan.recv_counts[r])]
            i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)
            pr, sc, rc = i32(peers), i32([plan.send_counts[r] for r in peers]), i32([plan.recv_counts[r] for r in peers])
            si = i32(np.concatenate([plan.send_index[r] for r in peers]) if peers else [])
            rs = i32(np.concatenate([plan.recv_index[r] for r in peers]) if peers else [])
            cached = plan._ghost_map_arrays = (len(peers), pr, sc, si, rc, rs)
        n, pr, sc, si, rc, rs = cached
        self._check(self.lib.eph_b200_set_ghost_map(self.h, n, pr.ctypes.data, sc.ctypes.data, si.ctypes.data,
                                                    rc.ctypes.data, rs.ctypes.data))
        self.exchange_bytes = 32 * (len(si) + len(rs))

    def exchange_ghosts(self):
        self._check(self.lib.eph_b200_exchange_ghosts(self.h))

    def set_grid_sharding(self, on=True):
        self._check(self.lib.eph_b200_set_grid_sharding(self.h, int(bool(on))))

    def reduce_and_solve(self, want_energy=True):
        e = C.c_double()
        self._check(self.lib.eph_b200_reduce_and_solve(self.h, C.byref(e) if want_energy else None))
        return e.value if want_energy else None

    def bind_grid_source(self, tensor):
        self._check(self.lib.eph_b200_bind_grid_source(self.h, tensor.data_ptr() if tensor is not None else None))
        self._src = tensor

    def set_grid_stream(self, stream):
        self._check(self.lib.eph_b200_set_grid_stream(self.h, C.c_void_p(stream) if stream else None))

    def set_comm_stream(self, stream):
        self._check(self.lib.eph_b200_set_comm_stream(self.h, C.c_void_p(stream) if stream else None))

    def set_boundary_atoms(self, index):
        """index: int32 local indices (numpy or device tensor) of the owned atoms other ranks hold as ghosts"""
        p = _ptr(index)
        self._check(self.lib.eph_b200_set_boundary_atoms(self.h, len(index), p[0], p[1]))

    def end_of_step(self, x, v, want_energy=True):
        ps = [_ptr(x), _ptr(v)]
        e = C.c_double()
        self._check(self.lib.eph_b200_end_of_step(self.h, ps[0][0], ps[1][0], C.byref(e) if want_energy else None,
                                                  _space(*ps)))
        return e.value if want_energy else None

    def initial_integrate(self, x, v, f, mass_by_type, dtv, dtf):
        m = np.ascontiguousarray(mass_by_type, dtype=np.float64)
        ps = [_ptr(x), _ptr(v), _ptr(f)]
        self._check(self.lib.eph_b200_initial_integrate(self.h, ps[0][0], ps[1][0], ps[2][0], m.ctypes.data, dtv, dtf, _space(*ps)))

    def final_integrate(self, v, f, mass_by_type, dtf):
        m = np.ascontiguousarray(mass_by_type, dtype=np.float64)
        ps = [_ptr(v), _ptr(f)]
        self._check(self.lib.eph_b200_final_integrate(self.h, ps[0][0], ps[1][0], m.ctypes.data, dtf, _space(*ps)))

    def refresh_ghosts(self, x, v):
        """ghost x, v follow their owners on the device (device tensors); the first call after set_atoms records the shifts"""
        self._check(self.lib.eph_b200_refresh_ghosts(self.h, x.data_ptr(), v.data_ptr()))

    # device-resident integration: host arrays in and out, x / v / f stay on the device between the hooks
    def resident_upload(self, x, v):
        self._check(self.lib.eph_b200_resident_upload(self.h, x.ctypes.data, v.ctypes.data))

    def resident_initial_integrate(self, f, mass_by_type, dtv, dtf, x_out, start_post_force_step=-1):
        m = np.ascontiguousarray(mass_by_type, dtype=np.float64)
        self._check(self.lib.eph_b200_resident_initial_integrate(self.h, None if f is None else f.ctypes.data, m.ctypes.data, dtv, dtf,
                                                                 None if x_out is None else x_out.ctypes.data, start_post_force_step))

    def resident_post_force(self, f, xi=None, step=0, f_out=None):
        self._check(self.lib.eph_b200_resident_post_force(self.h, f.ctypes.data, None if f_out is None else f_out.ctypes.data,
                                                          None if xi is None else xi.ctypes.data, step))

    def resident_final_integrate(self, mass_by_type, dtf, v_out=None):
        m = np.ascontiguousarray(mass_by_type, dtype=np.float64)
        self._check(self.lib.eph_b200_resident_final_integrate(self.h, m.ctypes.data, dtf, None if v_out is None else v_out.ctypes.data))

    def resident_get(self, which, out=None):
        if out is None:
            out = np.empty((self.nlocal, 3))
        self._check(self.lib.eph_b200_resident_get(self.h, which, out.ctypes.data))
        return out

    def resident_end_of_step(self, want_energy=True):
        e = C.c_double()
        self._check(self.lib.eph_b200_resident_end_of_step(self.h, C.byref(e) if want_energy else None))
        return e.value if want_energy else None

    def peratom(self):
        out = np.empty((self.nlocal, 8), dtype=np.float64)
        self._check(self.lib.eph_b200_get_peratom(self.h, out.ctypes.data, HOST))
        return out

    def probe(self, which):
        n = self.nlocal + self.nghost if which == 0 else 3 * self.nlocal
        out = np.empty(n, dtype=np.float64)
        self._check(self.lib.eph_b200_get_probe(self.h, which, out.ctypes.data))
        return out if which == 0 else out.reshape(-1, 3)

    def set_profiling(self, on=True):
        self._check(self.lib.eph_b200_set_profiling(self.h, int(on)))

    def kernel_times(self):
        """{kernel: (total ms, launches)} measured with CUDA events on the launch stream since set_profiling(True)"""
        names = (C.c_char_p * 32)()
        ms = (C.c_double * 32)()
        cnt = (C.c_longlong * 32)()
        n = self.lib.eph_b200_kernel_times(self.h, 32, names, ms, cnt)
        return {names[i].decode(): (ms[i], cnt[i]) for i in range(n)}

    def synchronize(self):
        self._check(self.lib.eph_b200_synchronize(self.h))

    def launch_count(self):
        return self.lib.eph_b200_launch_count(self.h)

    def status_word(self):
        out = C.c_uint()
        self._check(self.lib.eph_b200_status_word(self.h, C.byref(out)))
        return out.value
