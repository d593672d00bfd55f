"""Cost of the ordered (bit-reproducible) mode against the default kick, 1 M / 63^3 and 12.5 M / 127^3, one GPU."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ocelot_b200 import native
from bench import device_bunch, E_GEV, DZ

for n, mesh in ((1_000_000, 63), (12_500_000, 127)):
    p = device_bunch(torch, n, 1234, torch.device("cuda", 0))
    r, q = p.rparticles, p.q_array
    for ordered in (False, True):
        s = native.Solver(0, (mesh,) * 3)
        s.set_deterministic(ordered)
        s.enable_timers(True) if hasattr(s, "enable_timers") else None
        for _ in range(3):
            s.kick_device(r, q, E_GEV, DZ)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10):
            s.kick_device(r, q, E_GEV, DZ)
        b.record(); torch.cuda.synchronize()
        t = s.timers() if hasattr(s, "timers") else None
        print(f"n={n} mesh={mesh}^3 ordered={ordered}: {a.elapsed_time(b) / 10 * 1e3:.1f} us per kick (no graph when ordered / timers on); stages {t}")
        del s
