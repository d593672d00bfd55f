"""Particle-sharded space-charge kick over ``torch.distributed``.

One process per GPU; rank r owns a contiguous range of the bunch's particles
(SURVEY.md section 8e).  Per kick the ranks exchange only

  1. sum of Cartesian momenta + particle count      4 doubles, SUM   (sc.py:224)
  2. extents of the rotated, stretched coordinates  6 doubles, MAX
     and the charge centroid                         4 doubles, SUM
  3. the deposited charge grid rho                   nx*ny*nz doubles, SUM (sc.py:193)

Exchanges 1 and 2 run INSIDE the sweep kernels over NVLink peer memory (the block that finishes the
grid reduction pushes the rank's partials into every peer's mailbox, waits for theirs and folds them
in rank order), exchange 3 is a kernel of in-switch multimem reductions bracketed by two one-warp
barrier kernels; NCCL (all-reduce / all-gather) is the fallback when no peer mapping exists.

After step 3 every rank holds the same rho and solves the Poisson problem redundantly ("small-mesh
mode"); no particle ever crosses a link.  For large meshes ("slab mode") step 3 becomes
a reduce-scatter of rho into x-slabs and the FFT passes are split across ranks:

  3'. reduce-scatter rho -> x-slab ; z,y passes ; transpose ; x pass (FFT * K_hat * IFFT) ;
      transpose ; inverse y,z passes ; all-gather phi

With peer / multicast mappings none of these is a collective call: the transposes are the store side of
the y and x passes (output written straight into the consumer rank's buffer over NVLink), the inverse z
pass broadcasts its phi slab through the switch (multimem.st); without the mappings NCCL's all-to-all /
all-gather do the same job.

Every exchange runs in place on the native handle's device buffers, stream-ordered with the kernels, so
a kick involves no host synchronisation.

``StageEngine`` is the seam used by the CPU (gloo) tests: the product engine is
the native CUDA solver; tests substitute an engine backed by the oracle to check
the sharding / collective logic without a GPU.
"""
from __future__ import annotations

import numpy as np


def shard_bounds(n_total: int, world_size: int, rank: int):
    """Contiguous particle range [lo, hi) owned by ``rank`` (sizes differ by at most one)."""
    base, rem = divmod(int(n_total), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class NativeStageEngine:
    """The five stages of the kick on one GPU (include/ocelot_sc.h, ocl_sc_stage_*)."""

    fused_transposes = True     # slab mode: transposes as NVLink stores of the FFT passes (False: NCCL all-to-all)

    def __init__(self, device: int, nmesh_xyz, slab=None):
        """``slab`` = (rank, world) switches the Poisson solve to the slab-decomposed form."""
        from . import native
        self._native = native
        self.solver = native.Solver(device, nmesh_xyz)
        self.slab = slab
        if slab is not None:
            self.solver.slab_init(*slab)
        self.buffers = {
            "momentum": self.solver.collective_buffer(native.BUF_MOMENTUM),
            "extent_max": self.solver.collective_buffer(native.BUF_EXTENT_MAX),
            "extent_sum": self.solver.collective_buffer(native.BUF_EXTENT_SUM),
            "rho": self.solver.collective_buffer(native.BUF_RHO),
            "extent": self.solver.collective_buffer(native.BUF_EXTENT),
        }
        if slab is not None:
            for name, which in (("rho_slab", native.BUF_RHO_SLAB), ("xchg_a", native.BUF_XCHG_A),
                                ("xchg_b", native.BUF_XCHG_B), ("phi_slab", native.BUF_PHI_SLAB),
                                ("phi", native.BUF_PHI)):
                self.buffers[name] = self.solver.collective_buffer(which)
        self._gathered = None
        self.mailbox = None
        self.peer_rho = None
        self.nvls = None
        self.peer_xchg = None
        self.mc_phi = None
        self._draws = None
        self._E = None

    def setup_mailbox(self, dist, group, peer_rho=True, nvls=False):
        """Map one small symmetric buffer per rank into every rank (torch symmetric memory over
        NVLink peer access) and hand the peer pointers to the native handle: the two scalar
        exchanges then run as one tiny kernel each instead of an NCCL collective."""
        import torch
        import torch.distributed._symmetric_memory as symm
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        if world > 8:
            return False
        box = symm.empty(self.solver.MAILBOX_DOUBLES, dtype=torch.float64, device=self.buffers["rho"].device)
        box.zero_()
        hdl = symm.rendezvous(box, group=group if group is not None else dist.group.WORLD)
        torch.cuda.synchronize()
        dist.barrier(group=group)
        self.solver.mailbox_init(rank, world, list(hdl.buffer_ptrs))
        self.solver.defer_finish(False)        # the sweeps' finishing blocks now exchange over NVLink themselves
        self.mailbox = (box, hdl)              # keep the mapping alive
        if self.slab is not None and self.fused_transposes:
            # slab mode: the exchange buffers of the two transposes live in symmetric memory, and the y / x passes
            # store their output straight into the consumer rank's buffer over NVLink (no all-to-all)
            grp = group if group is not None else dist.group.WORLD
            ta = symm.empty(self.buffers["xchg_a"].numel(), dtype=torch.float64, device=box.device)
            tb = symm.empty(self.buffers["xchg_b"].numel(), dtype=torch.float64, device=box.device)
            ha, hb = symm.rendezvous(ta, group=grp), symm.rendezvous(tb, group=grp)
            torch.cuda.synchronize()
            dist.barrier(group=group)
            self.solver.set_peer_xchg(rank, world, list(ha.buffer_ptrs), list(hb.buffer_ptrs))
            torch.cuda.synchronize()
            dist.barrier(group=group)          # every rank has zeroed its buffers before anyone stores into them
            native = self._native
            self.buffers["xchg_a"] = self.solver.collective_buffer(native.BUF_XCHG_A)
            self.buffers["xchg_b"] = self.solver.collective_buffer(native.BUF_XCHG_B)
            self.peer_xchg = (ta, ha, tb, hb)
            # ... and the potential grid in symmetric memory with a multicast mapping: the inverse z pass broadcasts
            # its slab through the switch (multimem.st), which replaces the all-gather of phi
            tp = symm.empty(self.buffers["phi"].numel(), dtype=torch.float64, device=box.device)
            hp = symm.rendezvous(tp, group=grp)
            mcp = int(getattr(hp, "multicast_ptr", 0) or 0)
            torch.cuda.synchronize()
            dist.barrier(group=group)
            if mcp:
                self.solver.set_multicast_phi(hp.buffer_ptrs[rank], mcp)
                torch.cuda.synchronize()
                dist.barrier(group=group)
                self.buffers["phi"] = self.solver.collective_buffer(native.BUF_PHI)
                self.mc_phi = (tp, hp)
        if nvls:
            # rho in symmetric memory that is also mapped as one multicast range: the NVSwitch sums the
            # ranks' grids (multimem.ld_reduce) inside ocl_sc_nvls_reduce_rho -- no NCCL on the grid
            native = self._native
            grid = symm.empty(self.buffers["rho"].numel(), dtype=torch.float64, device=box.device)
            grid.zero_()
            ghdl = symm.rendezvous(grid, group=group if group is not None else dist.group.WORLD)
            mc = int(getattr(ghdl, "multicast_ptr", 0) or 0)
            torch.cuda.synchronize()
            dist.barrier(group=group)
            if mc:
                self.solver.set_multicast_rho(ghdl.buffer_ptrs[rank], mc)
                self.buffers["rho"] = self.solver.collective_buffer(native.BUF_RHO)
                self.nvls = (grid, ghdl)
                return True
            peer_rho = False                   # no multicast mapping on this system: stay on NCCL for rho
        if peer_rho:
            # opt-in: rho in plain symmetric memory; the first FFT pass sums the ranks' grids while loading them
            # over NVLink (no all-reduce / reduce-scatter kernel)
            native = self._native
            grid = symm.empty(self.buffers["rho"].numel(), dtype=torch.float64, device=box.device)
            grid.zero_()
            ghdl = symm.rendezvous(grid, group=group if group is not None else dist.group.WORLD)
            torch.cuda.synchronize()
            dist.barrier(group=group)
            self.solver.set_peer_rho(rank, world, list(ghdl.buffer_ptrs))
            self.buffers["rho"] = self.solver.collective_buffer(native.BUF_RHO)
            self.peer_rho = (grid, ghdl)
        return True

    def solve_slab(self, dist, group, draws):
        """Slab-decomposed solve: rho (local partial sums, nx_pad planes) -> field table."""
        b, s = self.buffers, self.solver
        if self.nvls is not None:
            s.nvls_reduce_rho()                # barrier, in-switch reduce-scatter into this rank's x-slab, barrier
        elif self.peer_rho is not None:
            s.mailbox_exchange(2)              # barrier: every rank's deposit is complete
        else:
            dist.reduce_scatter_tensor(b["rho_slab"], b["rho"], op=dist.ReduceOp.SUM, group=group)
        if self.peer_xchg is not None:
            s.slab_forward()                   # z, y passes; y output stored into the peers' x-pass input; barrier
            s.slab_xpass()                     # x pass; output stored into the peers' inverse-y input; barrier
        else:
            s.slab_forward()
            dist.all_to_all_single(b["xchg_b"], b["xchg_a"], group=group)
            s.slab_xpass()
            dist.all_to_all_single(b["xchg_a"], b["xchg_b"], group=group)
        s.slab_inverse()                       # with a multicast phi: broadcasts the slab through the switch + barrier
        if self.mc_phi is None:
            dist.all_gather_into_tensor(b["phi"], b["phi_slab"], group=group)
        s.slab_finish(draws)

    def combine_extents(self, dist, group):
        """One all-gather of the 10 extent doubles + an on-device fold (instead of a MAX and a
        SUM all-reduce)."""
        import torch
        world = dist.get_world_size(group)
        if self._gathered is None or self._gathered.numel() != 10 * world:
            self._gathered = torch.empty(10 * world, dtype=torch.float64, device=self.buffers["extent"].device)
        dist.all_gather_into_tensor(self._gathered, self.buffers["extent"], group=group)
        self.solver.combine_extents(self._gathered, world)

    def begin_kick(self, draws, multi):
        """Per-kick state: the mesh draws (the extent stage derives the mesh) and, for a sharded kick without
        a peer-memory mailbox, the deferred form (the caller all-reduces, then ``finish_*``)."""
        self._draws = draws
        self.solver.defer_finish(bool(multi) and self.mailbox is None)

    def momentum(self, r, q, E):
        self._E = E
        self.solver.stage_momentum(r, E)       # with a mailbox: exchange + frame inside the kernel's tail

    def finish_momentum(self):
        self.solver.stage_finish(0, self._E, self._draws)

    def extent(self, r, q, E):
        self.solver.stage_extent(r, q, E, self._draws)

    def finish_extent(self):
        self.solver.stage_finish(1, self._E, self._draws)

    def deposit(self, r, q, E, draws):
        self.solver.stage_deposit(r, q, E, draws)

    def solve(self, draws):
        self.solver.stage_solve(draws)

    def kick(self, r, E, dz, draws):
        self.solver.stage_kick(r, E, dz, draws)


def sharded_kick(engine, r, q, E_GeV, dz, draws=None, group=None, dist=None):
    """One SpaceCharge.apply on this rank's shard (in place).  ``engine`` exposes the
    stage methods and a ``buffers`` dict of tensors that ``dist.all_reduce`` can reduce."""
    if dz == 0:
        return
    if dist is None:
        import torch.distributed as dist
    multi = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
    SUM, MAX = dist.ReduceOp.SUM, dist.ReduceOp.MAX
    b = engine.buffers
    # with a peer-memory mailbox the two scalar exchanges happen inside the sweep kernels (the block that
    # finishes the reduction pushes / folds over NVLink): no collective call and no extra launch here
    fused = multi and getattr(engine, "mailbox", None) is not None
    if hasattr(engine, "begin_kick"):
        engine.begin_kick(draws, multi)
    engine.momentum(r, q, E_GeV)
    if multi and not fused:
        dist.all_reduce(b["momentum"], op=SUM, group=group)
        if hasattr(engine, "finish_momentum"):
            engine.finish_momentum()
    engine.extent(r, q, E_GeV)
    if multi and not fused:
        if hasattr(engine, "combine_extents"):
            engine.combine_extents(dist, group)
        else:
            dist.all_reduce(b["extent_max"], op=MAX, group=group)
            dist.all_reduce(b["extent_sum"], op=SUM, group=group)
        if hasattr(engine, "finish_extent"):
            engine.finish_extent()
    engine.deposit(r, q, E_GeV, draws)
    if getattr(engine, "slab", None) is not None:
        engine.solve_slab(dist, group, draws)      # works for any world size >= 1
    else:
        if multi and getattr(engine, "nvls", None) is not None:
            engine.solver.nvls_reduce_rho()        # barrier, in-switch all-reduce (multimem), barrier
        elif multi and getattr(engine, "peer_rho", None) is not None:
            engine.solver.mailbox_exchange(2)      # barrier; the solve's first pass sums the peers' grids
        elif multi:
            dist.all_reduce(b["rho"], op=SUM, group=group)
        engine.solve(draws)
    engine.kick(r, E_GeV, dz, draws)


class ShardedSpaceCharge:
    """PhysProc-style wrapper for a particle-sharded bunch: same attributes as
    ``SpaceCharge``; ``apply`` takes this rank's ``DeviceParticleArray`` shard."""

    # Meshes with any axis >= this use the slab-decomposed solve by default.  Measured [B200]: with the exchanges
    # fused into the FFT passes the slab form wins from 127^3 up (100 M / 127^3 on 2 / 4 / 8 GPUs: 3.06 / 1.69 /
    # 1.008 ms against 3.17 / 1.74 / 1.075 ms redundant); at 63^3 the redundant solve is one library graph.
    SLAB_MIN_MESH = 100

    def __init__(self, step=1, nmesh_xyz=(63, 63, 63), random_mesh=False, group=None, slab=None):
        self.step = step
        self.nmesh_xyz = list(nmesh_xyz)
        self.random_mesh = random_mesh
        self.random_seed = 10
        self.group = group
        self.slab = slab          # None: automatic (by mesh size); True / False: forced
        self.p2p = True           # scalar exchanges through NVLink peer memory instead of NCCL
        # Opt-in: fuse the charge-grid reduction into the first FFT pass (peer loads of every rank's
        # grid).  Correct (bit-identical to the NCCL path at W = 2) but measured 40 us SLOWER than an
        # all-reduce at 1M / 63^3 on 2 x B200 (8-byte peer loads in a latency-bound 125-block
        # kernel), so the default is the in-switch reduction below.
        self.p2p_rho = False
        # Default: the charge grid is summed inside the NVSwitch by the library's own kernel
        # (multimem.ld_reduce / multimem.st on a multicast mapping of the ranks' grids); falls back to the
        # NCCL all-reduce / reduce-scatter when the system offers no multicast mapping.
        self.nvls_rho = True
        self.use_graph = True
        self._engine = None
        self._graph = None
        self._graph_key = None

    def prepare(self, lat):
        if self.random_seed is not None:
            np.random.seed(self.random_seed)     # every rank draws the same stream (sc.py:104-107)

    def apply(self, p_shard, dz):
        if dz == 0:
            return
        import torch
        r, q = p_shard.rparticles, p_shard.q_array
        key = tuple(int(v) for v in self.nmesh_xyz)
        if self._engine is None or self._engine.solver.nmesh != key:
            import torch.distributed as dist
            slab = None
            want = self.slab if self.slab is not None else max(key) >= self.SLAB_MIN_MESH
            if want and dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1:
                slab = (dist.get_rank(self.group), dist.get_world_size(self.group))
            self._engine = NativeStageEngine(r.device.index or 0, key, slab=slab)
            if self.p2p and dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1:
                try:
                    self._engine.setup_mailbox(dist, self.group, peer_rho=self.p2p_rho,
                                               nvls=self.nvls_rho and not self.p2p_rho)
                except Exception as exc:  # noqa: BLE001  (no symmetric memory: stay on NCCL)
                    import logging
                    logging.getLogger(__name__).warning("peer-memory mailbox unavailable (%s); using NCCL", exc)
            self._graph, self._graph_key = None, None
        draws = None
        if self.random_mesh:
            draws = (np.random.uniform(low=1, high=1.1), np.random.uniform(low=-0.5, high=0.5))
        E = float(p_shard.E)
        if not self.use_graph:
            sharded_kick(self._engine, r, q, E, float(dz), draws, self.group)
            return
        eng = self._engine
        if eng.mailbox is not None and eng.nvls is not None and eng.slab is None and r.shape[1] > 0:
            # every exchange of this kick is one of the library's own kernels (mailbox + in-switch rho reduction,
            # redundant solve): the handle's kick_device IS the sharded kick, in the library's own CUDA graph
            eng.solver.defer_finish(False)
            eng.solver.kick_device(r, q, E, float(dz), draws)
            return
        # The staged kick *and* its NCCL collectives are captured once per particle buffer into a
        # CUDA graph; E, dz and the mesh draws live in a device block refreshed before each replay.
        solver = self._engine.solver
        gkey = (r.data_ptr(), r.stride(0), q.data_ptr(), r.shape[1])
        if self._graph is None or self._graph_key != gkey:
            sharded_kick(self._engine, r, q, E, float(dz), draws, self.group)   # also warms NCCL up
            torch.cuda.synchronize()
            solver.set_kick_params(E, 0.0, draws)       # dz = 0 while capturing: replays are the real kicks
            solver.use_device_params(True)
            try:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    sharded_kick(self._engine, r, q, E, float(dz), draws, self.group)
            finally:
                solver.use_device_params(False)
            self._graph, self._graph_key = g, gkey
            return                                       # this call's kick was the direct one above
        solver.set_kick_params(E, float(dz), draws)
        self._graph.replay()

    def exchange_status(self) -> int:
        """0: every exchange that runs inside the library's kernels has completed so far; 1 / 2 / 3: a momentum /
        extent / barrier-type exchange (rho reduction, slab transposes, phi broadcast) gave up after its ~4 s time-out
        -- the kernels terminate rather than hang the job when a peer is gone, and every result since is invalid.
        Synchronises the device."""
        if self._engine is None:
            return 0
        import torch
        torch.cuda.synchronize()
        return self._engine.solver.mailbox_status(True)

    def check(self):
        """Raise if a fused exchange timed out (see ``exchange_status``).  Call it before results are consumed on
        the host; ``finalize`` logs the same condition as an error."""
        code = self.exchange_status()
        if code:
            raise RuntimeError("sharded space-charge kick: " + self._STATUS.get(code, "exchange") + " timed out "
                               "waiting for a peer rank; the particle data of this rank is invalid since that kick")

    _STATUS = {1: "momentum exchange", 2: "extent exchange", 3: "barrier (rho reduction / slab exchange)"}

    def release(self):
        """Drop the captured graph.  NCCL requires graphs that captured its collectives to be
        destroyed before the communicator: call this (or finalize) before destroy_process_group."""
        self._graph, self._graph_key = None, None

    def finalize(self, *a, **k):
        self.release()
        code = self.exchange_status()
        if code:
            import logging
            logging.getLogger(__name__).error("sharded space-charge kick: %s timed out waiting for a peer rank; the "
                                              "particle data of this rank is invalid since that kick",
                                              self._STATUS.get(code, "exchange"))

    def __del__(self):
        try:
            self.release()
        except Exception:  # noqa: BLE001
            pass


# ---------------------------------------------------------------------------
# particle-sharded longitudinal space charge
# ---------------------------------------------------------------------------
def combine_lsc_stats(per_rank):
    """Fold the sweep-A statistics of the ranks (rows of [n, mean, M2, min, max, sum q, sum x, sum y])
    into those of the whole bunch; mean and centred square sum by the pairwise update of Chan et al."""
    a = np.asarray(per_rank, dtype=np.float64).reshape(-1, 8)
    a = a[a[:, 0] > 0]
    n = a[:, 0].sum()
    mean = np.sum(a[:, 0] * a[:, 1]) / n
    m2 = np.sum(a[:, 2] + a[:, 0] * (a[:, 1] - mean) ** 2)
    return dict(n=n, mean_tau=mean, m2_tau=m2, min_tau=a[:, 3].min(), max_tau=a[:, 4].max(),
                sum_q=a[:, 5].sum(), sum_x=a[:, 6].sum(), sum_y=a[:, 7].sum())


class NativeLscEngine:
    """The stages of an LSC kick on one GPU (include/ocelot_sc.h, ocl_sc_lsc_*)."""

    def __init__(self, device: int):
        from . import native
        self._native = native
        self.solver = native.Solver(device, (4, 4, 4))

    def stats(self, r, q):
        import torch
        s = self.solver.lsc_stats(r, q)
        return torch.tensor([s[k] for k in self.solver.LSC_STAT_KEYS], dtype=torch.float64, device=r.device)

    def deposit(self, r, params):
        import torch
        self.solver.lsc_deposit(r, params)
        nat = self._native
        return (self.solver.collective_buffer(nat.BUF_LSC_BINS).view(torch.int64),     # exact integer counts
                self.solver.collective_buffer(nat.BUF_LSC_SLICE_MAX),
                self.solver.collective_buffer(nat.BUF_LSC_SLICE_SUM))

    def solve_kick(self, r, params):
        self.solver.lsc_solve_kick(r, params)


def sharded_lsc_kick(engine, lsc, r, q, E_GeV, dz, group=None, dist=None):
    """One LSC.apply on this rank's shard (in place).  Exchanges per kick: one all-gather of 8
    statistics, then all-reduces of the integer current histogram (SUM, exact) and of the 4 + 5
    slice scalars (MAX, SUM).  ``lsc`` supplies the host-side scalars (ocelot_b200.LSC)."""
    if dz < 1e-10:                                                 # sc.py:566-568
        return None
    import torch
    if dist is None:
        import torch.distributed as dist
    multi = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
    mine = engine.stats(r, q)
    if multi:
        world = dist.get_world_size(group)
        allst = torch.empty(8 * world, dtype=torch.float64, device=mine.device)
        dist.all_gather_into_tensor(allst, mine, group=group)
        stats = combine_lsc_stats(allst.cpu().numpy())
    else:
        stats = combine_lsc_stats(mine.cpu().numpy())
    params = lsc.kick_parameters(stats, float(E_GeV), float(dz))
    bins, smax, ssum = engine.deposit(r, params)
    if multi:
        dist.all_reduce(bins, op=dist.ReduceOp.SUM, group=group)
        dist.all_reduce(smax, op=dist.ReduceOp.MAX, group=group)
        dist.all_reduce(ssum, op=dist.ReduceOp.SUM, group=group)
    engine.solve_kick(r, params)
    return params


class ShardedLSC:
    """PhysProc-style wrapper of ``LSC`` for a particle-sharded bunch: same constructor and
    attributes; ``apply`` takes this rank's ``DeviceParticleArray`` shard."""

    def __init__(self, step=1, group=None, **kwargs):
        from .lsc import LSC
        self.lsc = LSC(step=step, **kwargs)        # host scalars; the sharded kick always uses the staged form
        self.group = group
        self._engine = None

    def __getattr__(self, name):                    # step, bounds, smooth_param, z0, ... live on the LSC
        if name in ("lsc", "group", "_engine"):
            raise AttributeError(name)
        return getattr(self.lsc, name)

    def __setattr__(self, name, value):
        if name in ("lsc", "group", "_engine"):
            object.__setattr__(self, name, value)
        else:
            setattr(self.lsc, name, value)

    def prepare(self, lat):
        self.lsc.prepare(lat)

    def apply(self, p_shard, dz):
        r = p_shard.rparticles
        if self._engine is None:
            self._engine = NativeLscEngine(r.device.index or 0)
        self.lsc.last_params = sharded_lsc_kick(self._engine, self.lsc, r, p_shard.q_array, float(p_shard.E),
                                                float(dz), self.group)

    def finalize(self, *a, **k):
        pass
