"""Time the device LSC kick (sweep A + host scalars + deposit/solve/kick) on a resident bunch."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from ocelot_b200 import LSC, ParticleArray, DeviceParticleArray  # noqa: E402
from oracle import sc_oracle as orc  # noqa: E402  (bunch generator only)

for n in (1_000_000, 12_500_000):
    np.random.seed(1)
    r, q, E = orc.gaussian_bunch(n, energy=0.13, charge=250e-12)
    p = ParticleArray(n)
    p.rparticles[:], p.q_array[:], p.E = r, q, E
    dev = DeviceParticleArray.from_host(p)
    for sp in (False, True):
        lsc = LSC(step_profile=sp)
        for _ in range(3):
            lsc.apply(dev, 0.1)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        K = 20
        for _ in range(K):
            lsc.apply(dev, 0.1)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / K
        s = lsc._solver(0)
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        st = s.lsc_stats(dev.rparticles, dev.q_array)
        prm = lsc.kick_parameters(st, E, 0.1)
        e0.record(); s.lsc_deposit(dev.rparticles, prm); e1.record(); s.lsc_solve_kick(dev.rparticles, prm); e2.record()
        torch.cuda.synchronize()
        print(f"n={n} step_profile={sp} nb={prm['nb']}: {dt*1e6:.1f} us/kick wall "
              f"(deposit {e0.elapsed_time(e1)*1e3:.1f} us, solve+kick {e1.elapsed_time(e2)*1e3:.1f} us) "
              f"= {n/dt:.3e} particle-kicks/s")
