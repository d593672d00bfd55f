"""Particle-sharded space-charge kick over ``torch.distributed``.

One process per GPU; rank r owns a contiguous range of the bunch's particles
(SURVEY.md section 8e).  Per kick the ranks exchange only

  1. sum of Cartesian momenta + particle count      4 doubles, SUM   (sc.py:224)
  2. extents of the rotated, stretched coordinates  6 doubles, MAX   (sc.py:173,181)
     and the charge centroid                         4 doubles, SUM   (sc.py:182)
  3. the deposited charge grid rho                   nx*ny*nz doubles, SUM (sc.py:193)

after which every rank holds the same rho and solves the Poisson problem
redundantly ("small-mesh mode"); no particle ever crosses a link.  The
collectives run in place on the native handle's device buffers, stream-ordered
with the kernels, so a kick involves no host synchronisation.

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

    def __init__(self, device: int, nmesh_xyz):
        from . import native
        self._native = native
        self.solver = native.Solver(device, nmesh_xyz)
        self.buffers = {
            "momentum": self.solver.collective_buffer(native.BUF_MOMENTUM),
            "extent_max": self.solver.collective_buffer(native.BUF_EXTENT_MAX),
            "extent_sum": self.solver.collective_buffer(native.BUF_EXTENT_SUM),
            "rho": self.solver.collective_buffer(native.BUF_RHO),
        }

    def momentum(self, r, q, E):
        self.solver.stage_momentum(r, E)

    def extent(self, r, q, E):
        self.solver.stage_extent(r, q, E)

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
    engine.momentum(r, q, E_GeV)
    if multi:
        dist.all_reduce(b["momentum"], op=SUM, group=group)
    engine.extent(r, q, E_GeV)
    if multi:
        dist.all_reduce(b["extent_max"], op=MAX, group=group)
        dist.all_reduce(b["extent_sum"], op=SUM, group=group)
    engine.deposit(r, q, E_GeV, draws)
    if multi:
        dist.all_reduce(b["rho"], op=SUM, group=group)
    engine.solve(draws)
    engine.kick(r, E_GeV, dz, draws)


class ShardedSpaceCharge:
    """PhysProc-style wrapper for a particle-sharded bunch: same attributes as
    ``SpaceCharge``; ``apply`` takes this rank's ``DeviceParticleArray`` shard."""

    def __init__(self, step=1, nmesh_xyz=(63, 63, 63), random_mesh=False, group=None):
        self.step = step
        self.nmesh_xyz = list(nmesh_xyz)
        self.random_mesh = random_mesh
        self.random_seed = 10
        self.group = group
        self._engine = None

    def prepare(self, lat):
        if self.random_seed is not None:
            np.random.seed(self.random_seed)     # every rank draws the same stream (sc.py:104-107)

    def apply(self, p_shard, dz):
        if dz == 0:
            return
        r = p_shard.rparticles
        key = tuple(int(v) for v in self.nmesh_xyz)
        if self._engine is None or self._engine.solver.nmesh != key:
            self._engine = NativeStageEngine(r.device.index or 0, key)
        draws = None
        if self.random_mesh:
            draws = (np.random.uniform(low=1, high=1.1), np.random.uniform(low=-0.5, high=0.5))
        sharded_kick(self._engine, r, p_shard.q_array, float(p_shard.E), float(dz), draws, self.group)

    def finalize(self, *a, **k):
        pass
