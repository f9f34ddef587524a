"""FixEPHB200 -- the drop-in host class itself, not the Python harness -- on several ranks.

Each gloo rank owns a brick of the box and runs the product's fix inside the LAMMPS stand-in, whose MPI and
Comm::forward_comm(Fix*) are routed through torch.distributed for this test (tests/lammps_shim/mpi.h, lammps_shim.h).
Both multi-rank transports of the fix are driven:
  comm lammps   ghost values through Comm::forward_comm(Fix*) and the fix's pack/unpack callbacks, the grid source term
                through MPI_Allreduce (reference: fix_eph.cpp:743-744, :863-871, eph_fdm.h:479-491)
  comm nccl     the engine's own data plane: the fix builds the ghost map (one forward comm of (owner rank, owner index),
                MPI_Alltoall(v) of the request lists) and hands it to eph_b200_set_ghost_map; post_force / end_of_step then
                exchange {rho, W} (and the injected xi) and all-reduce the source term over NCCL
and the per-atom forces, densities, per-atom output, grid temperatures and energies must equal the UNMODIFIED reference
fix (oracle/_ref) run on the whole box on one rank, at 1e-10.  The engine is the host build of the device sources
(tests/emul) with its stand-in NCCL over gloo; on a B200 the same classes run over the real ones.  Test infrastructure."""
import os
import subprocess

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from eph_harness import harness as H
from eph_b200 import host
from eph_harness import parallel as P

from test_multirank_cpu import _free_port
from test_multirank_emulated import EMUL, _swap_in_emulated_engine, gloo_transport

TOL = 1e-10
CELLS, DT, STEPS = 6, 1e-4, 3


def group_ghosts_by_owner_rank(s, rank):
    """LAMMPS receives the ghosts of one swap as one contiguous range; the harness' bricks list them in image order.
    Re-order the ghosts (own images first, then by owner rank) and relabel the neighbour list accordingly."""
    nl, ng = s["nlocal"], s["nghost"]
    owner_rank = P.owner_rank_of(s["x"][nl:], s["box"], s["grid"])
    key = np.where(owner_rank == rank, -1, owner_rank)
    order = np.argsort(key, kind="stable")
    new_of_old = np.empty(ng, dtype=np.int64)
    new_of_old[order] = np.arange(ng)
    out = dict(s)
    for k in ("x", "v", "f", "type", "mask", "tag"):
        a = np.asarray(s[k])
        out[k] = np.ascontiguousarray(np.concatenate([a[:nl], a[nl:][order]]))
    out["ghost_owner"] = np.ascontiguousarray(np.asarray(s["ghost_owner"])[order])
    ne = np.asarray(s["neigh"]).copy()
    g = ne >= nl
    ne[g] = nl + new_of_old[ne[g] - nl]
    out["neigh"] = np.ascontiguousarray(ne.astype(np.int32))
    return out, owner_rank[order]


def _worker(rank, world, port, q, beta, gshape, comm_mode, extra):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), EPH_B200_P2P_WINDOW_MB="8")   # small shared-memory windows
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        L = _swap_in_emulated_engine()
        keep = [gloo_transport(L), host.FixDriver.plug_mpi(L, dist)]
        grid = P.brick_grid(world)
        s, ghost_rank = group_ghosts_by_owner_rank(H.make_system(CELLS, brick=(rank, grid)), rank)
        nl = s["nlocal"]
        plan = P.ExchangePlan(s, rank, world, dist)
        # the swaps of LAMMPS' forward comm: one per rank involved (this one included: its own periodic images)
        swaps = []
        for r in range(world):
            slots = nl + np.nonzero(ghost_rank == r)[0]
            send = plan.self_owner[ghost_rank == rank] if r == rank else plan.send_index[r]
            if len(slots) or len(send):
                assert len(slots) == 0 or np.array_equal(slots, np.arange(slots[0], slots[0] + len(slots)))
                swaps.append((r, send, int(slots[0]) if len(slots) else nl, len(slots)))
        s["ghost_owner"] = np.full(s["nghost"], -1, dtype=np.int32)   # several ranks: the stand-in uses the swaps instead
        args = H.fix_args(7, beta, ["Ni"], grid=gshape, style="eph/b200", extra=["rng", "mars", "comm", comm_mode] + list(extra))
        drv = host.FixDriver(s, args, dt=DT, lib=L)
        drv.set_swaps(swaps, dist)
        out = []
        for step in range(1, STEPS + 1):
            xi = np.random.default_rng(1000 + step).normal(size=(4 * CELLS ** 3, 3))[s["tag"][:nl] - 1]   # by atom tag
            drv.set_step(step)
            x, v, f = drv.xvf()
            f[:] = 0.0
            drv.update(f=f)
            drv.set_xi(xi)
            drv.post_force()
            drv.end_of_step()
            out.append(dict(f=drv.xvf()[2][:nl].copy(), rho=drv.probe(0)[:nl].copy(), array=drv.array().copy(), T=drv.grid_T().copy(),
                            Ee=drv.compute_vector(0), Tmean=drv.compute_vector(1)))
        q.put((rank, s["tag"][:nl].copy(), out, drv.n_forward()))
        drv.close()
        del keep
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,gshape,comm_mode,extra", [(2, (3, 2, 2), "lammps", ()), (2, (3, 2, 2), "nccl", ()),
                                                          (4, (6, 6, 8), "nccl", ("grid", "sharded")), (3, (2, 2, 2), "lammps", ())])
def test_fix_b200_on_several_ranks_matches_whole_box_reference(ni_trunc_beta, ref, world, gshape, comm_mode, extra):
    subprocess.check_call(["make", "-C", EMUL, "libeph_b200_emul.so"], stdout=subprocess.DEVNULL)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q, ni_trunc_beta, gshape, comm_mode, extra)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=600) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(60)
    # the unmodified reference fix on the whole box, one rank
    whole = H.make_system(CELLS)
    nlw = whole["nlocal"]
    drv = ref.fix_driver(whole, H.fix_args(7, ni_trunc_beta, ["Ni"], grid=gshape), dt=DT)
    order = np.argsort(whole["tag"][:nlw])
    assert sum(len(r[1]) for r in res) == nlw
    for k in range(STEPS):
        xi = np.random.default_rng(1001 + k).normal(size=(nlw, 3))[whole["tag"][:nlw] - 1]
        drv.set_step(k + 1)
        x, v, f = drv.xvf()
        f[:] = 0.0
        drv.update(f=f)
        drv.set_xi(xi)
        drv.post_force()
        drv.end_of_step()
        f_ref, rho_ref, arr_ref, T_ref = drv.xvf()[2][:nlw], drv.probe(0)[:nlw], drv.array(), drv.grid_T()
        for rank, tags, out, nfwd in res:
            idx = order[np.searchsorted(whole["tag"][:nlw][order], tags)]
            assert H.error_metrics(out[k]["f"], f_ref[idx], floor=np.abs(f_ref).max()) < TOL, (rank, k)
            assert H.error_metrics(out[k]["rho"], rho_ref[idx]) < TOL, (rank, k)
            assert H.error_metrics(out[k]["array"], arr_ref[idx], floor=np.abs(arr_ref).max()) < TOL, (rank, k)
            assert H.error_metrics(out[k]["T"], T_ref) < TOL, (rank, k)     # the source term was summed over the ranks
            assert abs(out[k]["Ee"] - drv.compute_vector(0)) <= 1e-9 * abs(drv.compute_vector(0)), (rank, k)
            assert abs(out[k]["Tmean"] - drv.compute_vector(1)) <= TOL * drv.compute_vector(1)
            # comm lammps: XI, RHO, WI per step like the reference; comm nccl: only the owner map at every re-neighbouring
            assert nfwd == (3 if comm_mode == "lammps" else 1) * STEPS, nfwd   # (the stand-in re-neighbours every step)
