import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ocelot_b200 import SpaceCharge, ParticleArray
n = 1_000_000
sc = SpaceCharge(nmesh_xyz=[63, 63, 63]); sc.prepare(None)
for pinned in (True, False):
    hp = ParticleArray(0)
    hr = torch.empty((6, n), dtype=torch.float64); hq = torch.empty(n, dtype=torch.float64)
    if pinned: hr, hq = hr.pin_memory(), hq.pin_memory()
    g = torch.Generator().manual_seed(1)
    sig = [1e-4, 2e-5, 1e-4, 2e-5, 1e-3, 1e-4]
    for k in range(6): hr[k] = torch.randn(n, generator=g, dtype=torch.float64) * sig[k]
    hq.fill_(250e-12 / n)
    hp.rparticles, hp.q_array, hp.E = hr.numpy(), hq.numpy(), 0.13
    for _ in range(3): sc.apply(hp, 0.1)
    t0 = time.perf_counter()
    for _ in range(10): sc.apply(hp, 0.1)
    dt = (time.perf_counter() - t0) / 10
    print("pinned" if pinned else "pageable", "ms/kick", dt * 1e3, "GB/s", 104e6 / dt / 1e9, "kicks/s %.3e" % (n / dt))
# raw copy bandwidth for reference
a = torch.empty(48_000_000 // 8, dtype=torch.float64).pin_memory(); d = torch.empty_like(a, device="cuda")
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(10): d.copy_(a, non_blocking=True)
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 10
print("H2D 48MB pinned GB/s", 48e6 / dt / 1e9)
t0 = time.perf_counter()
for _ in range(10): a.copy_(d, non_blocking=True)
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 10
print("D2H 48MB pinned GB/s", 48e6 / dt / 1e9)
