"""ctypes binding of the product's host-side C++ (libeph_b200_fix.so): `.beta`
table construction, FDM grid files and the FixEPHB200 class driven through the
LAMMPS stand-in.  `FixDriver` is generic over the driver prefix so the test
suite can drive the compiled reference fix with the very same calls."""
import ctypes as C

import numpy as np

from ._paths import lib_path

_fix = None


def load_fix_lib():
    global _fix
    if _fix is None:
        from . import lib as _engine
        _engine.load()  # libeph_b200_fix.so links libeph_b200.so
        _fix = C.CDLL(lib_path("fix"))
        _fix.ephh_last_error.restype = C.c_char_p
        for n in ("ephh_beta_load", "ephh_beta_from_knots", "ephh_grid_load"):
            getattr(_fix, n).restype = C.c_void_p
        _fix.ephh_grid_tables.restype = C.c_double
    return _fix


class BetaTables:
    """Spline tables built from a `.beta` file by the product's host code (fix/eph_tables.h)."""

    def __init__(self, path=None, knots=None, lib=None, prefix="ephh"):
        self.lib = lib or load_fix_lib()
        self.p = prefix
        if path is not None:
            f = getattr(self.lib, prefix + "_beta_load")
            f.restype = C.c_void_p
            self.h = f(str(path).encode())
        else:
            n_el, n_rho, dr, n_beta, drho, rc, rho_k, beta_k = knots
            rho_k = np.ascontiguousarray(rho_k, dtype=np.float64)
            beta_k = np.ascontiguousarray(beta_k, dtype=np.float64)
            self.h = self.lib.ephh_beta_from_knots(n_el, C.c_longlong(n_rho), C.c_double(dr), C.c_longlong(n_beta),
                                                   C.c_double(drho), C.c_double(rc), C.c_void_p(rho_k.ctypes.data),
                                                   C.c_void_p(beta_k.ctypes.data))
        if not self.h:
            raise RuntimeError("cannot build beta tables from %r" % (path,))
        dims = (C.c_longlong * 3)()
        scal = (C.c_double * 6)()
        getattr(self.lib, prefix + "_beta_info")(C.c_void_p(self.h), dims, scal)
        self.n_elements, self.n_rho, self.n_beta = (int(d) for d in dims)
        (self.r_cutoff, self.r_cutoff_sq, self.rho_cutoff, self.inv_dr, self.inv_dr_sq, self.inv_drho) = (float(s) for s in scal)

    def name(self, e):
        buf = C.create_string_buffer(64)
        getattr(self.lib, self.p + "_beta_name")(C.c_void_p(self.h), e, buf, 64)
        return buf.value.decode()

    def table(self, kind, element=None):
        """kind: 0 rho(r) 1 rho(r^2) 2 alpha 3 beta -> [n_elements][n][4] (or [n][4] for one element)"""
        n = self.n_rho if kind < 2 else self.n_beta
        els = range(self.n_elements) if element is None else [element]
        out = np.empty((len(els), n, 4), dtype=np.float64)
        for k, e in enumerate(els):
            getattr(self.lib, self.p + "_beta_table")(C.c_void_p(self.h), kind, e, C.c_void_p(out[k].ctypes.data))
        return out if element is None else out[0]

    def __del__(self):
        try:
            if self.h:
                getattr(self.lib, self.p + "_beta_free")(C.c_void_p(self.h))
        except Exception:
            pass


def spline_build(dx, y, lib=None, prefix="ephh"):
    lib = lib or load_fix_lib()
    y = np.ascontiguousarray(y, dtype=np.float64)
    out = np.empty((len(y), 4), dtype=np.float64)
    getattr(lib, prefix + "_spline_build")(C.c_double(dx), C.c_void_p(y.ctypes.data), len(y), C.c_void_p(out.ctypes.data))
    return out


class GridFile:
    """An FDM grid file parsed by the product's host code (fix/eph_grid_io.h)."""

    def __init__(self, path):
        self.lib = load_fix_lib()
        self.h = self.lib.ephh_grid_load(str(path).encode())
        if not self.h:
            raise RuntimeError(self.lib.ephh_last_error().decode())
        d = (C.c_longlong * 5)()
        box = (C.c_double * 6)()
        self.lib.ephh_grid_dims(C.c_void_p(self.h), d, box)
        self.nx, self.ny, self.nz, self.steps, self.n_T = (int(v) for v in d)
        self.box = np.array(list(box))
        self.ncell = self.nx * self.ny * self.nz

    def field(self, which):
        out = np.empty(self.ncell, dtype=np.float64)
        self.lib.ephh_grid_field(C.c_void_p(self.h), which, C.c_void_p(out.ctypes.data))
        return out

    def flags(self):
        fl = np.empty(self.ncell, dtype=np.int16)
        td = np.empty(self.ncell, dtype=np.uint16)
        self.lib.ephh_grid_flags(C.c_void_p(self.h), C.c_void_p(fl.ctypes.data), C.c_void_p(td.ctypes.data))
        return fl, td

    def tables(self):
        Ct = np.empty((self.n_T, 4)); Kt = np.empty((self.n_T, 4)); E = np.empty(self.n_T)
        dT = self.lib.ephh_grid_tables(C.c_void_p(self.h), C.c_void_p(Ct.ctypes.data), C.c_void_p(Kt.ctypes.data),
                                       C.c_void_p(E.ctypes.data))
        return dT, Ct, Kt, E

    def apply(self, engine):
        fl, td = self.flags()
        engine.set_grid(self.nx, self.ny, self.nz, self.box, self.field(0), self.field(2), self.field(3), self.field(4),
                        S_e=self.field(1), flag=fl, t_dyn=td, steps=self.steps)
        if self.n_T:
            engine.set_grid_tables(*self.tables())

    def write_heat_map(self, T, name, counter):
        T = np.ascontiguousarray(T, dtype=np.float64)
        if self.lib.ephh_grid_write_heat_map(C.c_void_p(self.h), C.c_void_p(T.ctypes.data), str(name).encode(), counter):
            raise RuntimeError(self.lib.ephh_last_error().decode())

    def write_restart(self, T, path):
        T = np.ascontiguousarray(T, dtype=np.float64)
        if self.lib.ephh_grid_write_restart(C.c_void_p(self.h), C.c_void_p(T.ctypes.data), str(path).encode()):
            raise RuntimeError(self.lib.ephh_last_error().decode())

    def __del__(self):
        try:
            if self.h:
                self.lib.ephh_grid_free(C.c_void_p(self.h))
        except Exception:
            pass


class FixError(RuntimeError):
    pass


class FixDriver:
    """A `fix eph`-style fix living in the LAMMPS stand-in (tests/lammps_shim/fix_driver.h).

    prefix "b200" drives FixEPHB200 (this library); the test suite passes the
    reference library with prefix "ref" to drive the unmodified FixEPH."""

    def __init__(self, system, fix_args, dt=1e-4, lib=None, prefix="b200", mass=None, neigh_modify=None):
        self.lib = lib or load_fix_lib()
        self.p = prefix
        self.sys = system
        self._fn("world_new").restype = C.c_void_p
        self._fn("last_error").restype = C.c_char_p
        self._fn("compute_vector").restype = C.c_double
        self._fn("grid_size").restype = C.c_longlong
        self._fn("n_forward").restype = C.c_longlong
        self._fn("neigh_cutoff").restype = C.c_double
        lo = np.zeros(3)
        hi = np.asarray(system["box"], dtype=np.float64)
        ntypes = int(system.get("ntypes", int(np.max(system["type"]))))
        m = np.ascontiguousarray(mass if mass is not None else [58.71] * ntypes, dtype=np.float64)
        self.w = self._fn("world_new")(C.c_longlong(system["natoms"]), ntypes, C.c_void_p(lo.ctypes.data),
                                       C.c_void_p(hi.ctypes.data), C.c_double(dt), C.c_void_p(m.ctypes.data))
        self.nlocal, self.nghost = system["nlocal"], system["nghost"]
        self._set_atoms(system)
        if neigh_modify is not None:          # (every, delay, check): a `neigh_modify` command ahead of the fix's init()
            self.neigh_modify(*neigh_modify)
        args = [str(a).encode() for a in fix_args]
        arr = (C.c_char_p * len(args))(*args)
        self._ck(self._fn("make_fix")(C.c_void_p(self.w), len(args), arr))
        self.set_neighbors(system["offsets"], system["neigh"])

    def _fn(self, name):
        return getattr(self.lib, self.p + "_" + name)

    def _ck(self, rc):
        if rc != 0:
            raise FixError(self._fn("last_error")(C.c_void_p(self.w)).decode())

    def _set_atoms(self, s):
        f = s.get("f")
        arrs = [np.ascontiguousarray(s["x"], dtype=np.float64), np.ascontiguousarray(s["v"], dtype=np.float64),
                None if f is None else np.ascontiguousarray(f, dtype=np.float64),
                np.ascontiguousarray(s["type"], dtype=np.int32), np.ascontiguousarray(s["mask"], dtype=np.int32),
                np.ascontiguousarray(s["tag"], dtype=np.int64), np.ascontiguousarray(s["ghost_owner"], dtype=np.int32)]
        self._ck(self._fn("set_atoms")(C.c_void_p(self.w), s["nlocal"], s["nghost"],
                                       *[C.c_void_p(a.ctypes.data) if a is not None else None for a in arrs]))

    # -- several ranks (tests): the stand-in's MPI and Comm::forward_comm(Fix*) get a transport ------------------
    MPI_ALLREDUCE = C.CFUNCTYPE(None, C.c_void_p, C.c_int, C.c_int, C.c_int)
    MPI_BCAST = C.CFUNCTYPE(None, C.c_void_p, C.c_int, C.c_int)
    MPI_ALLTOALLV = C.CFUNCTYPE(None, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int),
                                C.POINTER(C.c_int), C.POINTER(C.c_int))
    MPI_BARRIER = C.CFUNCTYPE(None)
    EXCHANGE = C.CFUNCTYPE(None, C.c_int, C.POINTER(C.c_int), C.POINTER(C.POINTER(C.c_double)), C.POINTER(C.c_int),
                           C.POINTER(C.POINTER(C.c_double)), C.POINTER(C.c_int))

    @classmethod
    def plug_mpi(cls, lib, dist, prefix="b200"):
        """Route the MPI stand-in of `lib` (tests/lammps_shim/mpi.h) through torch.distributed `dist`.  Call before the fix
        is constructed; returns the callbacks, which the caller keeps alive."""
        import torch
        rank, world = dist.get_rank(), dist.get_world_size()

        def view(ptr, n, ctype, dtype):
            return torch.from_numpy(np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ctype)), shape=(n,)).view(dtype))

        def allreduce(buf, n, dtype, op):
            t = view(buf, n, C.c_double if dtype == 1 else C.c_int, np.float64 if dtype == 1 else np.int32)
            dist.all_reduce(t, op={1: dist.ReduceOp.SUM, 2: dist.ReduceOp.MAX, 3: dist.ReduceOp.MIN}[op])

        def bcast(buf, nbytes, root):
            dist.broadcast(view(buf, nbytes, C.c_uint8, np.uint8), src=root)

        def alltoallv(send, scount, sdisp, recv, rcount, rdisp):
            outs = [torch.from_numpy(np.array([send[sdisp[r] + k] for k in range(scount[r])], dtype=np.int32)) for r in range(world)]
            ins = [torch.empty(rcount[r], dtype=torch.int32) for r in range(world)]
            ops = [dist.P2POp(dist.isend, outs[r], r) for r in range(world) if r != rank and scount[r]]
            ops += [dist.P2POp(dist.irecv, ins[r], r) for r in range(world) if r != rank and rcount[r]]
            if ops:
                for req in dist.batch_isend_irecv(ops):
                    req.wait()
            ins[rank] = outs[rank]
            for r in range(world):
                for k in range(rcount[r]):
                    recv[rdisp[r] + k] = int(ins[r][k])

        def barrier():
            dist.barrier()

        cbs = (cls.MPI_ALLREDUCE(allreduce), cls.MPI_BCAST(bcast), cls.MPI_ALLTOALLV(alltoallv), cls.MPI_BARRIER(barrier))
        getattr(lib, prefix + "_set_mpi")(rank, world, *cbs)
        return cbs

    def set_swaps(self, swaps, dist):
        """swaps: [(peer, sendlist (local indices, the peer's ghost order), first ghost index, count)], this rank included for
        its own periodic images.  Comm::forward_comm(Fix*) of the stand-in then moves the packed buffers over `dist`."""
        import torch

        def exchange(nswaps, peer, sbuf, sn, rbuf, rn):
            ops = []
            for k in range(nswaps):
                if sn[k]:
                    ops.append(dist.P2POp(dist.isend, torch.from_numpy(np.ctypeslib.as_array(sbuf[k], shape=(sn[k],))), peer[k]))
            for k in range(nswaps):
                if rn[k]:
                    ops.append(dist.P2POp(dist.irecv, torch.from_numpy(np.ctypeslib.as_array(rbuf[k], shape=(rn[k],))), peer[k]))
            if ops:
                for req in dist.batch_isend_irecv(ops):
                    req.wait()

        self._exchange_cb = self.EXCHANGE(exchange)
        i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)
        peer, cnt = i32([w[0] for w in swaps]), i32([len(w[1]) for w in swaps])
        flat = i32(np.concatenate([np.asarray(w[1], dtype=np.int32) for w in swaps]) if swaps else [])
        first, n = i32([w[2] for w in swaps]), i32([w[3] for w in swaps])
        self._ck(self._fn("set_swaps")(C.c_void_p(self.w), len(swaps), C.c_void_p(peer.ctypes.data), C.c_void_p(cnt.ctypes.data),
                                       C.c_void_p(flat.ctypes.data), C.c_void_p(first.ctypes.data), C.c_void_p(n.ctypes.data),
                                       self._exchange_cb))

    def set_neighbors(self, offsets, neigh):
        o = np.ascontiguousarray(offsets, dtype=np.int64)
        n = np.ascontiguousarray(neigh, dtype=np.int32)
        self._ck(self._fn("set_neighbors")(C.c_void_p(self.w), self.nlocal, C.c_void_p(o.ctypes.data), C.c_void_p(n.ctypes.data)))

    def update(self, x=None, v=None, f=None):
        a = [None if t is None else np.ascontiguousarray(t, dtype=np.float64) for t in (x, v, f)]
        self._ck(self._fn("update_xvf")(C.c_void_p(self.w), *[None if t is None else C.c_void_p(t.ctypes.data) for t in a]))

    def permute(self, new_of_old):
        """re-order the local atoms like LAMMPS' spatial sort does (atom i moves to new_of_old[i]); the fix is told through
        copy_arrays, the harness' arrays, ghost owners and neighbour list are relabelled"""
        p = np.ascontiguousarray(new_of_old, dtype=np.int32)
        assert len(p) == self.nlocal
        self._ck(self._fn("permute")(C.c_void_p(self.w), C.c_void_p(p.ctypes.data)))

    def set_xi(self, xi):
        xi = np.ascontiguousarray(xi, dtype=np.float64)
        self._ck(self._fn("set_xi")(C.c_void_p(self.w), C.c_void_p(xi.ctypes.data)))

    def set_dt(self, dt):
        self._ck(self._fn("set_dt")(C.c_void_p(self.w), C.c_double(dt)))

    def neigh_tick(self):
        """a step in which LAMMPS does not re-neighbour: the list ages (Neighbor::decide)"""
        self._fn("neigh_tick")(C.c_void_p(self.w))

    def neigh_modify(self, every=1, delay=0, check=True):
        """LAMMPS' `neigh_modify every N delay M check yes|no` in the stand-in"""
        self._fn("neigh_modify")(C.c_void_p(self.w), int(every), int(delay), int(bool(check)))

    def set_step(self, step):
        self._fn("set_step")(C.c_void_p(self.w), C.c_longlong(step))

    def initial_integrate(self): self._ck(self._fn("initial_integrate")(C.c_void_p(self.w)))
    def post_force(self): self._ck(self._fn("post_force")(C.c_void_p(self.w)))
    def final_integrate(self): self._ck(self._fn("final_integrate")(C.c_void_p(self.w)))
    def end_of_step(self): self._ck(self._fn("end_of_step")(C.c_void_p(self.w)))
    def post_run(self): self._ck(self._fn("post_run")(C.c_void_p(self.w)))
    def setmask(self): return self._fn("setmask")(C.c_void_p(self.w))
    def compute_vector(self, i): return self._fn("compute_vector")(C.c_void_p(self.w), i)
    def n_forward(self): return self._fn("n_forward")(C.c_void_p(self.w))
    def neigh_cutoff(self): return self._fn("neigh_cutoff")(C.c_void_p(self.w))

    def fix_flags(self):
        out = (C.c_int * 11)()
        self._fn("fix_flags")(C.c_void_p(self.w), out)
        names = ["vector_flag", "size_vector", "global_freq", "extvector", "nevery", "peratom_flag", "size_peratom_cols",
                 "peratom_freq", "comm_forward", "time_integrate", "ghost_velocity"]
        return dict(zip(names, list(out)))

    def xvf(self):
        n = self.nlocal + self.nghost
        x, v, f = (np.empty((n, 3)) for _ in range(3))
        self._fn("get_xvf")(C.c_void_p(self.w), C.c_void_p(x.ctypes.data), C.c_void_p(v.ctypes.data), C.c_void_p(f.ctypes.data))
        return x, v, f

    def array(self):
        out = np.empty((self.nlocal, self.fix_flags()["size_peratom_cols"]))
        self._fn("get_array")(C.c_void_p(self.w), C.c_void_p(out.ctypes.data))
        return out

    def probe(self, which, vec3=None):
        """0 rho[nt] 1 w 2 xi 3 f_EPH 4 f_RNG [nl][3]; `fix eph/atomic` adds 5 rho_a[nt] 6 E_a[nt] 7 dE_a[nl] 8 T_a[nl];
        `fix eph/coloured/exp` adds 5 f_dis 6 f_sto [nl][3] (pass vec3=True)"""
        nt = self.nlocal + self.nghost
        if vec3 is None:
            vec3 = which in (1, 2, 3, 4)
        n = 3 * self.nlocal if vec3 else nt if which in (0, 5, 6) else self.nlocal
        out = np.empty(n)
        self._ck(self._fn("get_probe")(C.c_void_p(self.w), which, C.c_void_p(out.ctypes.data)))
        return out.reshape(-1, 3) if vec3 else out

    def set_energy(self, E):
        """`fix eph/atomic` only: overwrite the per-atom electronic energies of the local atoms"""
        E = np.ascontiguousarray(E, dtype=np.float64)
        assert len(E) == self.nlocal
        self._ck(self._fn("set_energy")(C.c_void_p(self.w), C.c_void_p(E.ctypes.data)))

    def grid_T(self):
        out = np.empty(self._fn("grid_size")(C.c_void_p(self.w)))
        self._ck(self._fn("grid_T")(C.c_void_p(self.w), C.c_void_p(out.ctypes.data)))
        return out

    def close(self):
        if getattr(self, "w", None):
            self._fn("world_free")(C.c_void_p(self.w))
            self.w = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
