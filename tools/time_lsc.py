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
    for sp, ag in ((False, True), (False, False), (True, True)):
        lsc = LSC(step_profile=sp, async_grid=ag)
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
        print(f"n={n} step_profile={sp} device_grid={ag} nb={lsc.last_params['nb']}: {dt*1e6:.1f} us/kick wall "
              f"= {n/dt:.3e} particle-kicks/s")
