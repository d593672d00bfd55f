"""Particle-sharded kick on real GPUs: NCCL world of 2 (when the box has two GPUs)
must reproduce the single-GPU kick; with one GPU the sharded wrapper must equal
the plain plugin class."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from oracle import sc_oracle as orc  # noqa: E402


def _bunch(n, seed):
    np.random.seed(seed)
    r, q, E = orc.gaussian_bunch(n, energy=0.13, charge=250e-12)
    q = q * (0.5 + np.random.rand(n))
    return r, q, E


def _row_err(a, b):
    return max(float(np.max(np.abs(a[k] - b[k])) / np.std(b[k])) for k in range(6))


def test_sharded_wrapper_single_rank_equals_plugin():
    from ocelot_b200 import SpaceCharge, ParticleArray, DeviceParticleArray
    from ocelot_b200.distributed import ShardedSpaceCharge
    r0, q0, E = _bunch(100_000, 4)
    host = ParticleArray(r0.shape[1])
    host.rparticles[:], host.q_array[:], host.E = r0, q0, E
    a, b = DeviceParticleArray.from_host(host), DeviceParticleArray.from_host(host)
    sc = SpaceCharge(nmesh_xyz=[31, 31, 31])
    sh = ShardedSpaceCharge(nmesh_xyz=[31, 31, 31])
    sc.prepare(None); sh.prepare(None)
    sc.apply(a, 0.1)
    sh.apply(b, 0.1)
    assert _row_err(b.to_host().rparticles, a.to_host().rparticles) < 1e-12
    sh.check()                                 # failure detection: no fused exchange timed out (none exist at world 1)
    sh.finalize()
    assert sh.exchange_status() == 0


def _worker(rank, world, port, n, nmesh, out, slab=False, p2p_rho=False, nvls=False, p2p=True, empty_rank=None):
    import torch.distributed as dist
    from ocelot_b200 import ParticleArray, DeviceParticleArray
    from ocelot_b200.distributed import ShardedSpaceCharge, shard_bounds
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        r0, q0, E = _bunch(n, 4)
        lo, hi = shard_bounds(n, world, rank)
        if empty_rank is not None:             # one rank lost all its particles (e.g. to an aperture)
            lo, hi = (0, 0) if rank == empty_rank else (0, n)
        host = ParticleArray(hi - lo)
        host.rparticles[:], host.q_array[:], host.E = r0[:, lo:hi], q0[lo:hi], E
        shard = DeviceParticleArray.from_host(host, device=f"cuda:{rank}")
        sc = ShardedSpaceCharge(nmesh_xyz=list(nmesh), slab=slab)
        sc.p2p = p2p                           # False: no peer-memory mailbox, NCCL collectives + deferred finish
        sc.p2p_rho = p2p_rho
        sc.nvls_rho = nvls
        sc.prepare(None)
        for _ in range(2):
            sc.apply(shard, 0.1)
        for _ in range(2):                     # capture, replay
            sc.apply(shard, 0.1)
        torch.cuda.synchronize()
        out[rank] = (lo, hi, shard.to_host().rparticles.copy(), sc._engine.nvls is not None)
        sc.check()                             # no fused exchange gave up waiting for the peer
        sc.finalize()                          # graphs holding NCCL work must go before the communicator
    finally:
        torch.cuda.synchronize()
        dist.destroy_process_group()


@pytest.mark.parametrize("slab,p2p_rho,nvls,p2p,empty", [
    (False, False, False, True, None), (True, False, False, True, None), (False, True, False, True, None),
    (True, True, False, True, None), (False, False, True, True, None), (True, False, True, True, None),
    (False, False, False, False, None),        # pure NCCL: all-reduce / all-gather, then ocl_sc_stage_finish
    (True, False, False, False, None),
    (False, False, True, True, 1),             # rank 1 holds no particle: it still joins every exchange
    (False, False, False, False, 0)])
def test_two_gpu_sharded_kick_matches_single_gpu(slab, p2p_rho, nvls, p2p, empty):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    from ocelot_b200 import native
    n, nmesh, world = 400_001, (63, 63, 63), 2
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    with ctx.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, port, n, nmesh, out, slab, p2p_rho, nvls, p2p, empty), nprocs=world, join=True)
        parts = [out[k] for k in range(world)]
    r0, q0, E = _bunch(n, 4)
    solver = native.Solver(0, nmesh)
    r = torch.from_numpy(r0).cuda()
    q = torch.from_numpy(q0).cuda()
    for _ in range(4):
        solver.kick_device(r, q, E, 0.1)
    ref = r.cpu().numpy()
    got = np.empty_like(ref)
    for lo, hi, rr, used_nvls in parts:
        if hi > lo:
            got[:, lo:hi] = rr
        if nvls and not used_nvls:
            pytest.skip("no multicast mapping on this box: the NVLS reduction fell back to NCCL")
    # different reduction order across ranks: agreement to summation round-off (and its
    # amplification through the Green's function when the mesh step moves by an ulp)
    assert _row_err(got, ref) < 1e-10


# ---------------------------------------------------------------------------
# particle-sharded longitudinal space charge
# ---------------------------------------------------------------------------
def test_sharded_lsc_single_rank_equals_plugin():
    from ocelot_b200 import LSC, ParticleArray, DeviceParticleArray
    from ocelot_b200.distributed import ShardedLSC
    r0, q0, E = _bunch(100_000, 4)
    host = ParticleArray(r0.shape[1])
    host.rparticles[:], host.q_array[:], host.E = r0, q0, E
    a, b = DeviceParticleArray.from_host(host), DeviceParticleArray.from_host(host)
    LSC(step_profile=True).apply(a, 0.1)
    sh = ShardedLSC(step_profile=True)
    assert sh.step_profile is True and sh.bounds == [-0.4, 0.4]
    sh.apply(b, 0.1)
    assert np.array_equal(b.to_host().rparticles, a.to_host().rparticles)    # integer deposit: same bits


def _lsc_worker(rank, world, port, n, out):
    import torch.distributed as dist
    from ocelot_b200 import ParticleArray, DeviceParticleArray
    from ocelot_b200.distributed import ShardedLSC, shard_bounds
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        r0, q0, E = _bunch(n, 4)
        lo, hi = shard_bounds(n, world, rank)
        host = ParticleArray(hi - lo)
        host.rparticles[:], host.q_array[:], host.E = r0[:, lo:hi], q0[lo:hi], E
        shard = DeviceParticleArray.from_host(host, device=f"cuda:{rank}")
        lsc = ShardedLSC()
        for _ in range(3):
            lsc.apply(shard, 0.1)
        torch.cuda.synchronize()
        out[rank] = (lo, hi, shard.to_host().rparticles.copy())
    finally:
        torch.cuda.synchronize()
        dist.destroy_process_group()


def test_two_gpu_sharded_lsc_matches_single_gpu():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    from ocelot_b200 import LSC, ParticleArray, DeviceParticleArray
    n, world = 400_001, 2
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    with ctx.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_lsc_worker, args=(world, port, n, out), nprocs=world, join=True)
        parts = [out[k] for k in range(world)]
    r0, q0, E = _bunch(n, 4)
    host = ParticleArray(n)
    host.rparticles[:], host.q_array[:], host.E = r0, q0, E
    dev = DeviceParticleArray.from_host(host)
    lsc = LSC()
    for _ in range(3):
        lsc.apply(dev, 0.1)
    ref = dev.to_host().rparticles
    got = np.empty_like(ref)
    for lo, hi, rr in parts:
        got[:, lo:hi] = rr
    d_ref = ref[5] - r0[5]
    # the histogram is summed exactly; only the 8 + 9 floating-point statistics differ in rounding
    assert np.abs((got[5] - r0[5]) - d_ref).max() <= 1e-11 * np.abs(d_ref).max()
    assert np.array_equal(got[:5], ref[:5])
