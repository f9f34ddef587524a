"""Multi-GPU plumbing: LAMMPS-style spatial bricks, one rank per GPU, torch.distributed for transport.

Per step the path needs exactly one ghost exchange ({rho, Wx, Wy, Wz} from owners to ghosts, between the two
halves of post_force) and one all-reduce of the grid source term (between the two halves of end_of_step); xi needs
no exchange because ghosts regenerate the owner's Gaussians from the atom tag (include/eph_b200.h)."""
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


class GhostExchange:
    """Device-side exchange for an Engine: pack -> all_to_all over NCCL -> unpack.

    comm_stream: a torch.cuda.Stream registered with engine.set_comm_stream(); pack, the all-to-all and unpack then
    run on it, behind the boundary tiles of the density pass (engine.set_boundary_atoms(self.send_idx)), and overlap
    the sweep of the interior tiles on the engine's main stream."""

    def __init__(self, plan, dist, device, comm_stream=None):
        import torch
        self.plan, self.dist = plan, dist
        self.comm_stream = comm_stream
        self.send_idx = torch.as_tensor(plan.flat_send_index(), dtype=torch.int32, device=device)
        self.recv_idx = torch.as_tensor(plan.flat_recv_index(), dtype=torch.int32, device=device)
        self.send_buf = torch.empty((max(self.send_idx.numel(), 1), 4), dtype=torch.float64, device=device)
        self.recv_buf = torch.empty((max(self.recv_idx.numel(), 1), 4), dtype=torch.float64, device=device)
        self.out_splits = [4 * c for c in plan.send_counts]
        self.in_splits = [4 * c for c in plan.recv_counts]

    def __call__(self, engine):
        ns, nr = self.send_idx.numel(), self.recv_idx.numel()
        if ns:
            engine.pack_ghost_payload(self.send_idx, self.send_buf)
        if self.plan.world > 1:
            if self.comm_stream is not None:
                import torch
                with torch.cuda.stream(self.comm_stream):
                    self.dist.all_to_all_single(self.recv_buf.view(-1)[: 4 * nr], self.send_buf.view(-1)[: 4 * ns],
                                                output_split_sizes=self.in_splits, input_split_sizes=self.out_splits)
            else:
                self.dist.all_to_all_single(self.recv_buf.view(-1)[: 4 * nr], self.send_buf.view(-1)[: 4 * ns],
                                            output_split_sizes=self.in_splits, input_split_sizes=self.out_splits)
        if nr:
            engine.unpack_ghost_payload(self.recv_idx, self.recv_buf)

    def bytes_per_step(self):
        return 32 * (self.send_idx.numel() + self.recv_idx.numel())


def grid_slab(nz, rank, world):
    """z-planes [z0, z1) of the grid that `rank` updates in a sharded solve (equal slabs; None if nz does not divide)"""
    if world < 1 or nz % world:
        return None
    per = nz // world
    return rank * per, (rank + 1) * per


def sharded_grid_solve(engine, dist, rank, world):
    """EPH_FDM::solve (eph_fdm.h:267-400) with the grid sharded in z-slabs over the ranks: what the reference does as
    MPI_Allreduce + solve on rank 0 + MPI_Bcast (eph_fdm.h:481-491) becomes all-reduce (done by the caller) + slab
    sub-steps with one halo plane pair exchanged between sub-steps + all-gather of the slabs.

    engine: anything with grid_shape, grid_plan_substeps(), grid_substep(z0, z1) and grid_tensor(0) (the whole T_e
    field in the memory of this rank, current buffer; flat, z slowest).  Call between end_of_step_begin and
    end_of_step_end(external=True), on the stream the grid work is to run on.  Returns the number of sub-steps."""
    import torch
    nx, ny, nz = engine.grid_shape
    slab = grid_slab(nz, rank, world)
    if slab is None:
        raise ValueError("sharded grid solve needs nz (%d) divisible by the number of ranks (%d)" % (nz, world))
    z0, z1 = slab
    plane = nx * ny
    n = engine.grid_plan_substeps()
    prev, nxt = (rank - 1) % world, (rank + 1) % world
    for s in range(n):
        engine.grid_substep(z0, z1)
        if world == 1 or s == n - 1:
            continue
        # halo planes for the next sub-step, written in place at their global position (periodic in z).  Posting
        # order matters when prev == nxt (two ranks): sends go "bottom plane to prev, top plane to next", receives
        # "from next into the plane above the slab, from prev into the plane below", which pairs up on both sides.
        T = engine.grid_tensor(0).view(nz, plane)
        zlo, zhi = (z0 - 1) % nz, z1 % nz
        ops = [dist.P2POp(dist.isend, T[z0], prev), dist.P2POp(dist.isend, T[z1 - 1], nxt),
               dist.P2POp(dist.irecv, T[zhi], nxt), dist.P2POp(dist.irecv, T[zlo], prev)]
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    if world > 1 and n > 0:
        T = engine.grid_tensor(0).view(world, (z1 - z0) * plane)
        mine = T[rank].clone()
        if dist.get_backend() == "nccl":
            dist.all_gather_into_tensor(T.view(-1), mine)
        else:
            parts = [torch.empty_like(mine) for _ in range(world)]
            dist.all_gather(parts, mine)
            for r, p in enumerate(parts):
                T[r].copy_(p)
    return n


def distributed_step(engine, exchange, dist, x, v, f, step, dT_e, want_energy=False, grid_stream=None, sharded_grid=False):
    """One `fix eph` step on one rank of a multi-GPU run (device tensors).

    grid_stream: a torch.cuda.Stream registered with engine.set_grid_stream(); the all-reduce of the source term and
    the grid solve then run on it and overlap the next step's density pass.
    sharded_grid: every rank advances only its z-slab of the grid (sharded_grid_solve) instead of the whole grid."""
    engine.post_force_begin(x, v, None, step)
    exchange(engine)
    engine.post_force_end(f)
    engine.end_of_step_begin(None, v)   # positions are those of post_force (Verlet does not move atoms in between)
    multi = dist is not None and dist.get_world_size() > 1
    if multi and sharded_grid:
        import contextlib
        import torch
        with (torch.cuda.stream(grid_stream) if grid_stream is not None else contextlib.nullcontext()):
            dist.all_reduce(dT_e)
            sharded_grid_solve(engine, dist, dist.get_rank(), dist.get_world_size())
        return engine.end_of_step_end(want_energy, external=True)
    if multi:
        if grid_stream is not None:
            import torch
            with torch.cuda.stream(grid_stream):
                dist.all_reduce(dT_e)
        else:
            dist.all_reduce(dT_e)      # the reference's MPI_Allreduce of the grid source term (eph_fdm.h:481)
    return engine.end_of_step_end(want_energy)
