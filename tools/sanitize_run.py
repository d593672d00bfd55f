"""Small end-to-end exercise of every kernel for compute-sanitizer (memcheck / racecheck)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("OCL_SC_GRAPH", "0")
from ocelot_b200 import native, SpaceCharge, ParticleArray, DeviceParticleArray, apply_map, get_envelope
from ocelot_b200.beam import apply_cavity
from ocelot_b200 import RectAperture, EllipticalAperture, LSC
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
    # round 2: every gather mode (z-fastest / lane pairs staged / lane pairs with register prefetch), the deferred
    # finish form, the aperture compaction kernels, LSC in both grid modes
    for mode in ("0", "1", "2"):
        os.environ["OCL_SC_GATHER"] = mode
        sg = native.Solver(0, nm)
        sg.kick_device(r, q, dev.E, 0.1)
        sg.field_at_particles(r, q, dev.E)
    os.environ.pop("OCL_SC_GATHER")
    so = native.Solver(0, nm)                 # ordered (deterministic) deposit: cell index sweep, radix sort, ordered sums
    so.set_deterministic(True)
    so.kick_device(r, q, dev.E, 0.1)
    sd = native.Solver(0, nm)
    sd.defer_finish(True)
    sd.stage_momentum(r, dev.E); sd.stage_finish(0, dev.E); sd.stage_extent(r, q, dev.E); sd.stage_finish(1, dev.E)
    sd.stage_deposit(r, q, dev.E); sd.stage_solve(); sd.stage_kick(r, dev.E, 0.1)
    RectAperture(xmin=-1.5e-4, xmax=2e-4, ymax=1e-4).apply(dev, 0.0)
    EllipticalAperture(xmax=1.8e-4, ymax=0.9e-4, dx=1e-5).apply(dev, 0.0)
    LSC(async_grid=True).apply(dev, 0.1); LSC(async_grid=False).apply(dev, 0.1)
    sc.apply(dev, 0.1)
    torch.cuda.synchronize()
    print("ok", n, dev.n, nm, t)
