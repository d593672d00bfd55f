"""Probe (GPU): are the device mesh steps bit-identical to the oracle's, and what is the field error?"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ocelot_b200 import native
from oracle import sc_oracle as orc

def ulps(a, b):
    return [int(abs(np.float64(x).view(np.int64) - np.float64(y).view(np.int64))) for x, y in zip(a, b)]

for n, nm, seeds in ((1_000_000, 63, (5, 6, 7, 8, 11)), (2_000_000, 127, (9, 10)), (200_000, 31, (1, 2, 3))):
    for seed in seeds:
        np.random.seed(seed)
        r0, q0, E = orc.gaussian_bunch(n, energy=0.13, charge=250e-12)
        taps = {}
        orc.sc_kick(r0.copy(), q0, E, 0.1, (nm,) * 3, fft="padded", workers=8, taps=taps)
        s = native.Solver(0, (nm,) * 3)
        r, q = torch.from_numpy(r0).cuda(), torch.from_numpy(q0).cuda()
        Ex = s.field_at_particles(r, q, E).cpu().numpy()
        g = s.geometry()
        err = max(np.max(np.abs(Ex[:, c] - taps["Exyz"][:, c])) / np.max(np.abs(taps["Exyz"][:, c])) for c in range(3))
        print(n, nm, seed, "steps ulp diff", ulps(g["steps"], taps["steps"]), "gamma0 ulp", ulps([g["gamma0"]], [taps["gamma0"]]),
              "xoff rel", np.max(np.abs(g["X_off"] - taps["X_off"]) / np.abs(taps["X_off"])), "field err %.2e" % err, flush=True)
