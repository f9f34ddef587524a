"""The multi-GPU path of the product, end to end on the CPU: `world` gloo ranks, each with the host build of the engine
(tests/emul, see test_engine_emulated.py) and its brick of the box, run what bench.py runs on B200s: plain
eph_b200_post_force / eph_b200_end_of_step on an engine with a communicator attached (eph_b200_comm_init) and a ghost
map registered (eph_b200_set_ghost_map).  The ENGINE issues the ghost exchange (grouped ncclSend / ncclRecv of {rho, W}),
the all-reduce of the grid source term and, with grid sharding, the halo planes and the all-gather of the slab solve;
on the host build a stand-in NCCL (tests/emul/fake_nccl.cpp) hands the bytes to torch.distributed over gloo.  Forces,
densities, grid temperatures and energies must equal the single-rank oracle on the whole box at 1e-10.
Test infrastructure only."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from eph_harness import harness as H
from eph_b200 import host, lib
from eph_harness import parallel as P
from oracle import oracle as O

from test_multirank_cpu import _free_port

EMUL = os.path.join(os.path.dirname(os.path.abspath(__file__)), "emul")
TOL = 1e-10
CELLS, SEED, DT = 6, 4711, 1e-4

P2P_FN = C.CFUNCTYPE(None, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_void_p), C.POINTER(C.c_size_t))
ALLREDUCE_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_size_t)
ALLGATHER_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p, C.c_size_t)


def _bytes_view(ptr, n):
    return torch.from_numpy(np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), shape=(n,)))


def gloo_transport(L):
    """the stand-in NCCL of the host build moves its bytes through these callbacks: torch.distributed over gloo"""
    def p2p(nops, is_send, peer, buf, nbytes):
        ops = [dist.P2POp(dist.isend if is_send[k] else dist.irecv, _bytes_view(buf[k], nbytes[k]), peer[k]) for k in range(nops)]
        for req in dist.batch_isend_irecv(ops):
            req.wait()

    def allreduce(buf, count):
        t = torch.from_numpy(np.ctypeslib.as_array(C.cast(buf, C.POINTER(C.c_double)), shape=(count,)))
        dist.all_reduce(t)

    def allgather(send, recv, nbytes):
        world = dist.get_world_size()
        mine = _bytes_view(send, nbytes).clone()
        parts = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(parts, mine)
        out = _bytes_view(recv, nbytes * world)
        for r, p in enumerate(parts):
            out[r * nbytes:(r + 1) * nbytes] = p

    cbs = (P2P_FN(p2p), ALLREDUCE_FN(allreduce), ALLGATHER_FN(allgather))
    L.emul_nccl_set_transport(*cbs)
    L.emul_nccl_calls.restype = C.c_longlong
    return cbs   # keep alive


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
    return L


def _worker(rank, world, port, q, beta, gshape, sharded, boundary_first, tau0=0.0, inject_xi=False, transport="p2p"):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), EPH_B200_P2P_WINDOW_MB="8")
    if transport == "nccl":
        os.environ["EPH_B200_EXCHANGE"] = "nccl"
    else:
        os.environ.pop("EPH_B200_EXCHANGE", None)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        L = _swap_in_emulated_engine()
        keep_cbs = gloo_transport(L)
        grid = P.brick_grid(world)
        s = H.make_system(CELLS, brick=(rank, grid))
        nl = s["nlocal"]
        box = [0, s["box"][0], 0, s["box"][1], 0, s["box"][2]]
        plan = P.ExchangePlan(s, rank, world, dist)
        eng = lib.Engine([0], flags=7, seed=SEED, rank=rank, nranks=world)
        eng.set_tables_from(host.BetaTables(path=beta))
        eng.set_grid(*gshape, box, 300.0, 1.0, 3.5e-6, 0.1248)
        eng.set_dt(DT)
        if tau0 > 0:
            eng.set_colour(tau0)          # fix eph/coloured/exp: the memory kernel filters each rank's own atoms
        P.attach_comm(eng, dist, rank, world)
        eng.set_grid_sharding(sharded)
        if boundary_first:
            # the density pass sweeps the tiles other ranks wait for first (tile_mark / tile_split), and with a
            # communication stream registered the pack waits on the finished-tile counter (cuStreamWaitValue32) and
            # post_force_end on the unpack event; streams are immediate on the host build, so this checks the
            # bookkeeping of that path, not its overlap
            eng.set_comm_stream(0x10)
        keep = [np.ascontiguousarray(s["type"], dtype=np.int32), np.ascontiguousarray(s["mask"], dtype=np.int32),
                np.ascontiguousarray(s["tag"], dtype=np.int64), np.ascontiguousarray(plan.self_owner, dtype=np.int32),
                np.ascontiguousarray(s["offsets"], dtype=np.int64), np.ascontiguousarray(s["neigh"], dtype=np.int32)]
        eng.set_atoms(nl, s["nghost"], *keep[:4])
        eng.set_neighbors(*keep[4:])
        eng.set_ghost_map(plan)
        x, v = s["x"].copy(), s["v"].copy()
        f = np.zeros((nl, 3))
        out = []
        for step in (1, 2, 3):
            f[:] = 0.0
            xi = O.xi_stream(SEED + 17, step, s["tag"][:nl]) if inject_xi else None   # caller-supplied Gaussians: the XI exchange runs too
            E = P.distributed_step(eng, x, v, f, step, xi=xi, want_energy=True)
            out.append(dict(f=f.copy(), rho=eng.probe(0)[:nl].copy(), T=eng.get_grid(0).copy(), E=E,
                            substeps=eng.last_substeps()))
        calls = [L.emul_nccl_calls(k) for k in range(4)]
        # LAMMPS' forward comm of x and v on the "device" (eph_b200_refresh_ghosts, the second per-step exchange of a run
        # whose atoms live there): the first call records the image shifts, then the owners move and the ghosts must follow
        import torch
        tag = s["tag"].astype(np.float64)
        move = lambda t: 1e-3 * np.stack([np.sin(t), np.cos(2 * t), np.sin(3 * t)], axis=1)
        vel = lambda t: np.stack([1e-3 * t, -2e-3 * t, 5e-4 * t], axis=1)
        xt, vt = torch.from_numpy(s["x"].copy()), torch.from_numpy(s["v"].copy())
        eng.refresh_ghosts(xt, vt)
        xt[:nl] += torch.from_numpy(move(tag[:nl]))
        vt[:nl] = torch.from_numpy(vel(tag[:nl]))
        eng.refresh_ghosts(xt, vt)
        refresh_err = max(np.abs(xt.numpy()[nl:] - (s["x"][nl:] + move(tag[nl:]))).max(), np.abs(vt.numpy()[nl:] - vel(tag[nl:])).max())
        q.put((rank, s["tag"][:nl].copy(), out, eng.exchange_bytes, calls, eng.comm_transport(), float(refresh_err)))
        del keep_cbs
    finally:
        dist.destroy_process_group()


# transport of the two ghost exchanges: "p2p" = the peer-memory windows of csrc/eph_p2p.cuh (shared-memory objects between
# the ranks' processes on the host build: the same kernels, flags, epochs and double-buffered halves as over NVLink),
# "nccl" = grouped send / receive through the stand-in library
@pytest.mark.parametrize("world,gshape,sharded,boundary_first,tau0,inject_xi,transport",
                         [(2, (3, 2, 2), False, False, 0.0, False, "p2p"), (2, (3, 2, 2), False, True, 0.0, True, "p2p"),
                          (2, (8, 8, 8), True, False, 0.0, False, "nccl"), (4, (8, 8, 8), True, True, 0.0, False, "p2p"),
                          (3, (6, 6, 9), True, False, 0.0, True, "p2p"), (2, (3, 2, 2), False, True, 5e-4, False, "nccl"),
                          (4, (3, 2, 2), False, False, 0.0, True, "nccl")])
def test_engine_data_plane_on_gloo_ranks_matches_whole_box_oracle(synth_beta_1, world, gshape, sharded, boundary_first, tau0, inject_xi, transport):
    subprocess.check_call(["make", "-C", EMUL, "libeph_b200_emul.so"], stdout=subprocess.DEVNULL)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q, synth_beta_1, gshape, sharded, boundary_first, tau0, inject_xi, transport)) for r in range(world)]
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
    assert sum(len(r[1]) for r in res) == nlw
    for k, step in enumerate((1, 2, 3)):
        # the ranks generate the same Gaussians from the atom tags, or are handed them and exchange the ghosts' share
        xi = O.xi_stream(SEED + 17 if inject_xi else SEED, step, whole["tag"][:nlw])
        fx.f[:] = 0.0
        fx.post_force(xi)
        fx.end_of_step()
        ref_f, ref_rho = fx.f[:nlw], np.array(fx.ptr(0))[:nlw]
        E = 0.0
        for rank, tags, out, nbytes, calls, used, refresh_err in res:
            assert refresh_err < 1e-12, (rank, refresh_err)
            idx = order[np.searchsorted(whole["tag"][:nlw][order], tags)]
            assert H.error_metrics(out[k]["f"], ref_f[idx], floor=np.abs(ref_f).max()) < TOL, (rank, step)
            assert H.error_metrics(out[k]["rho"], ref_rho[idx]) < TOL, (rank, step)
            assert H.error_metrics(out[k]["T"], fx.fdm.field(0)) < TOL, (rank, step)
            assert nbytes > 0
            # the engine itself issued the collectives.  Per step: one all-reduce of the source term (or reduce-scatter +
            # all-gather), and with send / receive one grouped exchange; setting up the transport costs one all-gather (the
            # window handles) and one all-reduce (the agreement), every set_ghost_map over peer memory one more all-reduce
            assert used == (2 if transport == "p2p" else 1), used
            assert calls[2] == 3 + (2 if transport == "p2p" else 1) and calls[3] == 1 + (3 if sharded else 0), calls
            if transport == "nccl":
                assert calls[0] >= 3, calls
            elif not sharded:
                assert calls[0] == 0, calls     # no send / receive at all: the rows went through the windows
            if sharded:   # fine grid: the solve takes several sub-steps, so halo planes were exchanged between them
                assert out[k]["substeps"] >= 3
            E += out[k]["E"]
        assert abs(E - (fx.Ee() - E_prev)) < 1e-9 * abs(E)
        E_prev = fx.Ee()


def _small_window_worker(rank, world, port, q, beta):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), EPH_B200_P2P_WINDOW_MB="0.0625")
    os.environ.pop("EPH_B200_EXCHANGE", None)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        L = _swap_in_emulated_engine()
        keep_cbs = gloo_transport(L)
        grid = P.brick_grid(world)
        s = H.make_system(CELLS, brick=(rank, grid))
        plan = P.ExchangePlan(s, rank, world, dist)
        eng = lib.Engine([0], flags=7, seed=SEED, rank=rank, nranks=world)
        eng.set_tables_from(host.BetaTables(path=beta))
        P.attach_comm(eng, dist, rank, world)
        keep = [np.ascontiguousarray(s["type"], dtype=np.int32), np.ascontiguousarray(s["mask"], dtype=np.int32),
                np.ascontiguousarray(s["tag"], dtype=np.int64), np.ascontiguousarray(plan.self_owner, dtype=np.int32)]
        eng.set_atoms(s["nlocal"], s["nghost"], *keep)
        msg = None
        try:
            eng.set_ghost_map(plan)
        except lib.EphError as e:
            msg = str(e)
        q.put((rank, eng.comm_transport(), msg, int(max(max(plan.send_counts), max(plan.recv_counts)))))
        del keep_cbs
    finally:
        dist.destroy_process_group()


def test_ghost_rows_that_do_not_fit_the_window_fail_on_every_rank(synth_beta_1):
    """64 KB windows: the ghost rows of this box do not fit a rank's share.  set_ghost_map must say so with the size it
    needs -- on ALL ranks (the verdict travels with the collective of the registration), not leave the others waiting."""
    subprocess.check_call(["make", "-C", EMUL, "libeph_b200_emul.so"], stdout=subprocess.DEVNULL)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    world = 2
    procs = [ctx.Process(target=_small_window_worker, args=(r, world, port, q, synth_beta_1)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=300) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(60)
    for rank, used, msg, rows in res:
        assert used == 2, used                      # the windows themselves were set up
        assert rows * 104 > (65536 - 8192) // world // 2, rows   # the premise of the test
        assert msg is not None and "EPH_B200_P2P_WINDOW_MB" in msg and "EPH_B200_EXCHANGE=nccl" in msg, msg


def _timeout_worker(rank, world, port, q, beta):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), EPH_B200_P2P_WINDOW_MB="8", EPH_B200_P2P_TIMEOUT_MS="300")
    os.environ.pop("EPH_B200_EXCHANGE", None)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        L = _swap_in_emulated_engine()
        keep_cbs = gloo_transport(L)
        grid = P.brick_grid(world)
        s = H.make_system(CELLS, brick=(rank, grid))
        nl = s["nlocal"]
        box = [0, s["box"][0], 0, s["box"][1], 0, s["box"][2]]
        plan = P.ExchangePlan(s, rank, world, dist)
        eng = lib.Engine([0], flags=7, seed=SEED, rank=rank, nranks=world)
        eng.set_tables_from(host.BetaTables(path=beta))
        eng.set_grid(2, 2, 2, box, 300.0, 1.0, 3.5e-6, 0.1248)
        eng.set_dt(DT)
        P.attach_comm(eng, dist, rank, world)
        keep = [np.ascontiguousarray(s["type"], dtype=np.int32), np.ascontiguousarray(s["mask"], dtype=np.int32),
                np.ascontiguousarray(s["tag"], dtype=np.int64), np.ascontiguousarray(plan.self_owner, dtype=np.int32),
                np.ascontiguousarray(s["offsets"], dtype=np.int64), np.ascontiguousarray(s["neigh"], dtype=np.int32)]
        eng.set_atoms(nl, s["nghost"], *keep[:4])
        eng.set_neighbors(*keep[4:])
        eng.set_ghost_map(plan)
        x, v, f = s["x"].copy(), s["v"].copy(), np.zeros((nl, 3))
        msg = None
        try:
            if rank == 0:
                eng.post_force(x, v, f, None, 1)          # exchanges: rank 1 never sends its rows
            else:
                eng.post_force_begin(x, v, None, 1)       # a rank that skips the exchange (a bug, a crash ...)
                eng.post_force_end(f)
            eng.end_of_step(x, v, want_energy=True)       # the all-reduce of the source term still matches on both
        except lib.EphError as e:
            msg = str(e)
        q.put((rank, msg, eng.status_word()))
        del keep_cbs
    finally:
        dist.destroy_process_group()


def test_a_peer_that_never_delivers_is_an_error_not_a_hang(synth_beta_1):
    """rank 1 skips the ghost exchange of a step: rank 0's receiving kernel gives up after the time-out (300 ms here),
    sets bit 8 of the status word, and the step's host synchronisation in end_of_step turns it into an error"""
    subprocess.check_call(["make", "-C", EMUL, "libeph_b200_emul.so"], stdout=subprocess.DEVNULL)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_timeout_worker, args=(r, 2, port, q, synth_beta_1)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=300) for _ in range(2)], key=lambda t: t[0])
    for p in procs:
        p.join(60)
    (r0, msg0, st0), (r1, msg1, st1) = res
    assert msg0 is not None and "timed out" in msg0, msg0
    assert st0 & 0x100 and not (st1 & 0x100), (st0, st1)
    assert msg1 is None, msg1
