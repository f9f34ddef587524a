"""Multi-GPU harness: LAMMPS-style spatial bricks, one rank per GPU.

The data plane is NOT here: the ghost exchange of post_force, the all-reduce of the grid source term and the halo planes
of the sharded grid solve are issued by the engine itself over NCCL (include/eph_b200.h: eph_b200_comm_init,
set_ghost_map, and then plain post_force / end_of_step).  This module plays LAMMPS' part around it -- who owns which
atom (ExchangePlan, built from atom tags like the fix builds its ghost map through Comm::forward_comm) -- and uses
torch.distributed only to carry the 128-byte communicator id and the set-up lists between the ranks."""
import numpy as np


def brick_grid(nranks):
    """Processor grid like LAMMPS picks for a cubic box: as cubic as possible."""
    best = (nranks, 1, 1)
    for px in range(1, nranks + 1):
        if nranks % px:
            continue
        for py in range(1, nranks // px + 1):
            if (nranks // px) % py:
                continue
            pz = nranks // px // py
            cand = tuple(sorted((px, py, pz), reverse=True))
            if max(cand) - min(cand) < max(best) - min(best):
                best = cand
    return best


def owner_rank_of(x, box, grid):
    """Rank whose brick contains each (possibly shifted image) position."""
    L = np.asarray(box, dtype=np.float64)
    g = np.asarray(grid)
    xw = np.mod(x, L)
    xw[xw >= L] = 0.0
    c = np.minimum((xw / (L / g)).astype(np.int64), g - 1)
    return (c[:, 0] + g[0] * (c[:, 1] + g[1] * c[:, 2])).astype(np.int64)


class ExchangePlan:
    """Who sends which owned atoms to whom, built once per re-neighbouring from atom tags only.

    send_index[r]  local indices of my atoms that rank r holds as ghosts, in the order r asked for them
    recv_index[r]  my ghost slots (nlocal + g) that rank r fills, in the order I asked
    self_owner     ghost_owner array for set_atoms: local owner index for my own periodic images, -1 for remote ghosts
    """

    def __init__(self, system, rank, world, dist=None, asked=None):
        nl = system["nlocal"]
        tags = np.asarray(system["tag"])
        ghost_tags = tags[nl:]
        ghost_rank = owner_rank_of(np.asarray(system["x"])[nl:], system["box"], system["grid"])
        local_tags = tags[:nl]
        order = np.argsort(local_tags, kind="stable")
        sorted_tags = local_tags[order]

        def to_local(req):
            pos = np.searchsorted(sorted_tags, req)
            if len(req) and (np.any(pos >= nl) or np.any(sorted_tags[np.minimum(pos, nl - 1)] != req)):
                raise RuntimeError("exchange plan: asked for an atom this rank does not own")
            return order[pos].astype(np.int32)

        self.rank, self.world = rank, world
        self.self_owner = np.full(len(ghost_tags), -1, dtype=np.int32)
        mine = ghost_rank == rank
        self.self_owner[mine] = to_local(ghost_tags[mine])
        want = [ghost_tags[ghost_rank == r] if r != rank else ghost_tags[:0] for r in range(world)]
        self.recv_index = [(nl + np.nonzero(ghost_rank == r)[0]).astype(np.int32) if r != rank else np.zeros(0, np.int32)
                           for r in range(world)]
        self.want = want
        if asked is not None:
            pass                      # in-process construction (build_all)
        elif world == 1:
            asked = [want[0]]
        else:
            asked = _alltoall_lists(want, dist)
        self.send_index = [to_local(np.asarray(a, dtype=np.int64)) for a in asked]
        self.send_counts = [len(a) for a in self.send_index]
        self.recv_counts = [len(a) for a in self.recv_index]

    @classmethod
    def build_all(cls, systems):
        """Plans of all ranks inside one process (tests, single-process drivers): no communication needed."""
        world = len(systems)
        first = [cls(s, r, world, asked=[np.zeros(0, np.int64)] * world) for r, s in enumerate(systems)]
        return [cls(s, r, world, asked=[first[q].want[r] for q in range(world)]) for r, s in enumerate(systems)]

    def flat_send_index(self):
        return np.concatenate(self.send_index) if self.world > 0 else np.zeros(0, np.int32)

    def flat_recv_index(self):
        return np.concatenate(self.recv_index)


def _alltoall_lists(lists, dist):
    """Exchange variable-length int64 lists between all ranks (set-up time, any backend)."""
    import torch
    world = dist.get_world_size()
    backend = dist.get_backend()
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    counts = torch.tensor([len(a) for a in lists], dtype=torch.int64, device=dev)
    theirs = torch.empty(world, dtype=torch.int64, device=dev)
    dist.all_to_all_single(theirs, counts)
    out_splits = [int(c) for c in counts.tolist()]
    in_splits = [int(c) for c in theirs.tolist()]
    send = torch.as_tensor(np.concatenate(lists) if sum(out_splits) else np.zeros(0, np.int64), dtype=torch.int64, device=dev)
    recv = torch.empty(sum(in_splits), dtype=torch.int64, device=dev)
    dist.all_to_all_single(recv, send, output_split_sizes=in_splits, input_split_sizes=out_splits)
    recv = recv.cpu().numpy()
    return np.split(recv, np.cumsum(in_splits)[:-1])


def grid_slab(nz, rank, world):
    """z-planes [z0, z1) of the grid that `rank` advances in a sharded solve (equal slabs; None if nz does not divide)"""
    if world < 1 or nz % world:
        return None
    per = nz // world
    return rank * per, (rank + 1) * per


def attach_comm(engine, dist, rank, world):
    """Give the engine its NCCL communicator: rank 0 creates the id, torch.distributed carries the 128 bytes (what
    MPI_Bcast does in the fix), every rank joins (eph_b200_comm_init is collective)."""
    box = [engine.comm_get_id() if rank == 0 else None]
    if world > 1:
        dist.broadcast_object_list(box, src=0)
    engine.comm_init(box[0], rank, world)


def distributed_step(engine, x, v, f, step, xi=None, want_energy=False):
    """One `fix eph` step on one rank of a multi-GPU run: the same two calls as on one rank.  With a communicator attached
    and a ghost map registered the engine does the ghost exchange, the source all-reduce and the (sharded) grid solve
    itself.  Positions at end_of_step are those of post_force (Verlet does not move atoms in between)."""
    engine.post_force(x, v, f, xi, step)
    return engine.end_of_step(None, v, want_energy)
