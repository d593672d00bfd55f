"""Small end-to-end exercise of every kernel for compute-sanitizer (memcheck / racecheck)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("OCL_SC_GRAPH", "0")
from ocelot_b200 import native, SpaceCharge, ParticleArray, DeviceParticleArray, apply_map, get_envelope
from ocelot_b200.beam import apply_cavity
rng = np.random.RandomState(0)
for n, nm in ((3001, (9, 12, 7)), (20000, (31, 31, 31)), (5000, (20, 33, 64))):
    host = ParticleArray(n)
    sig = [1e-4, 2e-5, 1e-4, 2e-5, 1e-3, 1e-4]
    for k in range(6): host.rparticles[k] = rng.randn(n) * sig[k]
    host.q_array[:] = 1e-10 / n; host.E = 0.05
    dev = DeviceParticleArray.from_host(host)
    sc = SpaceCharge(nmesh_xyz=list(nm)); sc.prepare(None)
    sc.apply(dev, 0.1); sc.apply(host, 0.1)
    apply_map(dev, np.eye(6) + 0.01 * rng.randn(6, 6), rng.randn(6) * 1e-7, rng.randn(6, 6, 6) * (rng.rand(6, 6, 6) < 0.2))
    apply_cavity(dev, np.eye(6), None, 0.02, 18.0, 1.3e9, 0.1, 1.0)
    t = get_envelope(dev)
    s = native.Solver(0, nm)
    s.slab_init(0, 1)
    r, q = dev.rparticles, dev.q_array
    s.stage_momentum(r, dev.E); s.stage_extent(r, q, dev.E); s.stage_deposit(r, q, dev.E)
    s.collective_buffer(native.BUF_RHO_SLAB).copy_(s.collective_buffer(native.BUF_RHO)[:s.collective_buffer(native.BUF_RHO_SLAB).numel()])
    s.slab_forward(); s.collective_buffer(native.BUF_XCHG_B).copy_(s.collective_buffer(native.BUF_XCHG_A)); s.slab_xpass()
    s.collective_buffer(native.BUF_XCHG_A).copy_(s.collective_buffer(native.BUF_XCHG_B)); s.slab_inverse()
    s.collective_buffer(native.BUF_PHI).copy_(s.collective_buffer(native.BUF_PHI_SLAB)[:s.collective_buffer(native.BUF_PHI).numel()]) if s.collective_buffer(native.BUF_PHI).numel() <= s.collective_buffer(native.BUF_PHI_SLAB).numel() else None
    s.slab_finish(); s.stage_kick(r, dev.E, 0.1)
    torch.cuda.synchronize()
    print("ok", n, nm, t)
