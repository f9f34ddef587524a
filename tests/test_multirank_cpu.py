"""Host-side logic of the multi-GPU path on CPU: brick decomposition and the ghost exchange plan, with two
gloo ranks.  (The compute itself needs GPUs: tests/test_gpu_parity.py::test_two_rank_engine_matches_oracle.)"""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from eph_harness import harness as H
from eph_harness import parallel as P


def test_brick_grid_shapes():
    assert P.brick_grid(1) == (1, 1, 1) and P.brick_grid(2) == (2, 1, 1)
    assert P.brick_grid(4) == (2, 2, 1) and P.brick_grid(8) == (2, 2, 2)


def test_bricks_partition_the_box_and_ghost_shells_are_complete():
    grid = (2, 2, 1)
    whole = H.make_system(6)
    seen = []
    for r in range(4):
        s = H.make_system(6, brick=(r, grid))
        seen.append(s["tag"][: s["nlocal"]])
        # every list entry within the cut-off is present: compare one atom's in-cutoff neighbour tags with the whole box
        i = 3
        t = s["tag"][i]
        nb = np.sort(s["tag"][s["neigh"][s["offsets"][i]: s["offsets"][i + 1]]])
        iw = int(np.nonzero(whole["tag"][: whole["nlocal"]] == t)[0][0])
        nbw = np.sort(whole["tag"][whole["neigh"][whole["offsets"][iw]: whole["offsets"][iw + 1]]])
        assert np.array_equal(nb, nbw)
        assert np.array_equal(P.owner_rank_of(s["x"][: s["nlocal"]], s["box"], grid), np.full(s["nlocal"], r))
    alltags = np.sort(np.concatenate(seen))
    assert np.array_equal(alltags, np.arange(1, whole["natoms"] + 1))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q, cells=6):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        grid = P.brick_grid(world)
        s = H.make_system(cells, brick=(rank, grid))
        plan = P.ExchangePlan(s, rank, world, dist)
        nl, ng = s["nlocal"], s["nghost"]
        # payload = a function of the atom tag; after the exchange every ghost must hold its owner's payload
        pay = np.zeros((nl + ng, 4))
        pay[:nl] = np.stack([s["tag"][:nl] * 1.0, s["tag"][:nl] * 2.0, -s["tag"][:nl] * 1.0, np.sqrt(s["tag"][:nl])], axis=1)
        send = torch.as_tensor(pay[plan.flat_send_index()].reshape(-1))
        recv = torch.empty(4 * sum(plan.recv_counts), dtype=torch.float64)
        dist.all_to_all_single(recv, send, output_split_sizes=[4 * c for c in plan.recv_counts],
                               input_split_sizes=[4 * c for c in plan.send_counts])
        pay[plan.flat_recv_index()] = recv.numpy().reshape(-1, 4)
        own = plan.self_owner >= 0
        pay[nl:][own] = pay[plan.self_owner[own]]
        gt = s["tag"][nl:].astype(np.float64)
        ok = np.array_equal(pay[nl:], np.stack([gt, 2 * gt, -gt, np.sqrt(gt)], axis=1))
        # grid all-reduce: each rank deposits its own atoms, the sum must equal the whole-box deposit
        cell = np.minimum((s["x"][:nl] / (s["box"] / 4)).astype(int), 3)
        hist = np.zeros(64)
        np.add.at(hist, cell[:, 0] + 4 * cell[:, 1] + 16 * cell[:, 2], 1.0)
        t = torch.as_tensor(hist)
        dist.all_reduce(t)
        q.put((rank, bool(ok), float(t.sum()), int((~own).sum()), int(own.sum())))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("cells", [6, (10, 5, 5)])   # cubic box (strong scaling); 5^3 cells per rank (bench.py --weak)
def test_exchange_plan_two_gloo_ranks(cells):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q, cells)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(60)
    natoms = 4 * (cells ** 3 if np.isscalar(cells) else int(np.prod(cells)))
    for rank, ok, total, remote, own in res:
        assert ok, "ghost payload mismatch on rank %d" % rank
        assert total == natoms
        assert remote > 0 and own > 0     # 2 x 1 x 1 bricks: x-ghosts are remote, y/z images are the rank's own atoms


# The sharded grid solve itself (slab sub-steps, halo planes in place, all-gather) is driven by the engine
# (eph_b200_reduce_and_solve); tests/test_multirank_emulated.py and tests/test_multirank_fix.py run it on 2, 3 and 4 ranks.
def test_grid_slabs_partition_the_planes():
    assert P.grid_slab(64, 3, 8) == (24, 32) and P.grid_slab(6, 0, 1) == (0, 6)
    assert P.grid_slab(10, 0, 4) is None
    for world in (1, 2, 4, 8):
        planes = [z for r in range(world) for z in range(*P.grid_slab(64, r, world))]
        assert planes == list(range(64))
