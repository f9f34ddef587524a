"""The multi-GPU path of the product, end to end on the CPU: `world` gloo ranks, each with the host build of the engine
(tests/emul, see test_engine_emulated.py) and its brick of the box, run eph_b200.parallel.distributed_step -- the very
function bench.py runs over NCCL: post_force_begin, ghost payload all-to-all (GhostExchange), post_force_end, deposit,
all-reduce of the grid source term, replicated or sharded grid solve with halo planes -- and the forces, densities,
grid temperatures and energies must equal the single-rank oracle on the whole box at 1e-10.  On the host build device
memory is host memory, so CPU tensors stand where device tensors stand on a B200.  Test infrastructure only."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from eph_b200 import harness as H
from eph_b200 import host, lib
from eph_b200 import parallel as P
from oracle import oracle as O

from test_multirank_cpu import _free_port

EMUL = os.path.join(os.path.dirname(os.path.abspath(__file__)), "emul")
TOL = 1e-10
CELLS, SEED, DT = 6, 4711, 1e-4


def _swap_in_emulated_engine():
    L = C.CDLL(os.path.join(EMUL, "libeph_b200_emul%s.so" % os.environ.get("EPH_EMUL_SUFFIX", "")))
    for name, (res, args) in lib.SYMBOLS.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    L.ephh_last_error.restype = C.c_char_p
    for n in ("ephh_beta_load", "ephh_beta_from_knots", "ephh_grid_load"):
        getattr(L, n).restype = C.c_void_p
    L.ephh_grid_tables.restype = C.c_double
    lib._lib, host._fix = L, L

    def grid_tensor(self, which=0):   # a CPU view of the engine's field: device memory is host memory here
        p = C.c_void_p()
        self._check(self.lib.eph_b200_grid_device_ptr(self.h, which, C.byref(p)))
        return torch.from_numpy(np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_double)), shape=(self.ncell,)))

    lib.Engine.grid_tensor = grid_tensor


def _worker(rank, world, port, q, beta, gshape, sharded, boundary_first, tau0=0.0):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        _swap_in_emulated_engine()
        grid = P.brick_grid(world)
        s = H.make_system(CELLS, brick=(rank, grid))
        nl = s["nlocal"]
        box = [0, s["box"][0], 0, s["box"][1], 0, s["box"][2]]
        plan = P.ExchangePlan(s, rank, world, dist)
        exch = P.GhostExchange(plan, dist, torch.device("cpu"))
        eng = lib.Engine([0], flags=7, seed=SEED, rank=rank, nranks=world)
        eng.set_tables_from(host.BetaTables(path=beta))
        eng.set_grid(*gshape, box, 300.0, 1.0, 3.5e-6, 0.1248)
        eng.set_dt(DT)
        if tau0 > 0:
            eng.set_colour(tau0)          # fix eph/coloured/exp: the memory kernel filters each rank's own atoms
        keep = [np.ascontiguousarray(s["type"], dtype=np.int32), np.ascontiguousarray(s["mask"], dtype=np.int32),
                np.ascontiguousarray(s["tag"], dtype=np.int64), np.ascontiguousarray(plan.self_owner, dtype=np.int32),
                np.ascontiguousarray(s["offsets"], dtype=np.int64), np.ascontiguousarray(s["neigh"], dtype=np.int32)]
        eng.set_atoms(nl, s["nghost"], *keep[:4])
        eng.set_neighbors(*keep[4:])
        src = torch.zeros(int(np.prod(gshape)), dtype=torch.float64)
        eng.bind_grid_source(src)
        if boundary_first:
            # the density pass sweeps the tiles other ranks wait for first (tile_mark / tile_split), and with a
            # communication stream registered the pack waits on the finished-tile counter (cuStreamWaitValue32) and
            # post_force_end on the unpack event; streams are immediate on the host build, so this checks the
            # bookkeeping of that path, not its overlap
            eng.set_comm_stream(0x10)
            eng.set_boundary_atoms(np.ascontiguousarray(plan.flat_send_index(), dtype=np.int32))
        x, v = torch.as_tensor(s["x"].copy()), torch.as_tensor(s["v"].copy())
        f = torch.zeros((nl, 3), dtype=torch.float64)
        out = []
        for step in (1, 2):
            f.zero_()
            E = P.distributed_step(eng, exch, dist, x, v, f, step, src, want_energy=True, sharded_grid=sharded)
            out.append(dict(f=f.numpy().copy(), rho=eng.probe(0)[:nl].copy(), T=eng.get_grid(0).copy(), E=E,
                            substeps=eng.last_substeps()))
        q.put((rank, s["tag"][:nl].copy(), out, exch.bytes_per_step()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,gshape,sharded,boundary_first,tau0", [(2, (3, 2, 2), False, False, 0.0), (2, (3, 2, 2), False, True, 0.0),
                                                                      (2, (8, 8, 8), True, False, 0.0), (4, (8, 8, 8), True, True, 0.0),
                                                                      (2, (3, 2, 2), False, True, 5e-4)])
def test_distributed_step_on_gloo_ranks_matches_whole_box_oracle(synth_beta_1, world, gshape, sharded, boundary_first, tau0):
    subprocess.check_call(["make", "-C", EMUL, "libeph_b200_emul.so"], stdout=subprocess.DEVNULL)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q, synth_beta_1, gshape, sharded, boundary_first, tau0)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=600) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(60)
    whole = H.make_system(CELLS)
    nlw = whole["nlocal"]
    box = [0, whole["box"][0], 0, whole["box"][1], 0, whole["box"][2]]
    fx = O.Fix(whole, O.Beta(path=synth_beta_1), O.FDM(*gshape, box, 300.0, 3.5e-6, 1.0, 0.1248), 7, dt=DT)
    if tau0 > 0:
        fx.set_colour(tau0)
    order = np.argsort(whole["tag"][:nlw])
    E_prev = 0.0
    assert sum(len(tags) for _, tags, _, _ in res) == nlw
    for k, step in enumerate((1, 2)):
        xi = O.xi_stream(SEED, step, whole["tag"][:nlw])   # the ranks generate the same Gaussians from the atom tags
        fx.f[:] = 0.0
        fx.post_force(xi)
        fx.end_of_step()
        ref_f, ref_rho = fx.f[:nlw], np.array(fx.ptr(0))[:nlw]
        E = 0.0
        for rank, tags, out, nbytes in res:
            idx = order[np.searchsorted(whole["tag"][:nlw][order], tags)]
            assert H.error_metrics(out[k]["f"], ref_f[idx], floor=np.abs(ref_f).max()) < TOL, (rank, step)
            assert H.error_metrics(out[k]["rho"], ref_rho[idx]) < TOL, (rank, step)
            assert H.error_metrics(out[k]["T"], fx.fdm.field(0)) < TOL, (rank, step)
            assert nbytes > 0
            if sharded:   # fine grid: the solve takes several sub-steps, so halo planes were exchanged between them
                assert out[k]["substeps"] >= 3
            E += out[k]["E"]
        assert abs(E - (fx.Ee() - E_prev)) < 1e-9 * abs(E)
        E_prev = fx.Ee()
