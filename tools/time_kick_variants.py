"""Time one workload's stage breakdown (used for A/B runs of kernel variants)."""
import os, sys, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ocelot_b200 import native
def run(n, nm, kicks=30):
    g = torch.Generator(device="cuda").manual_seed(1)
    r = torch.empty((6, n), dtype=torch.float64, device="cuda")
    sig = [1e-4, 2e-5, 1e-4, 2e-5, 1e-3, 1e-4]
    for k in range(6):
        r[k] = torch.randn(n, generator=g, device="cuda", dtype=torch.float64) * sig[k]
    r[5] += 0.01 * r[4] / 1e-3
    q = torch.full((n,), 250e-12 / n, dtype=torch.float64, device="cuda")
    s = native.Solver(0, (nm,) * 3)
    for _ in range(5): s.kick_device(r, q, 0.13, 0.1)
    s.enable_timers(True); acc = {}
    for _ in range(kicks):
        s.kick_device(r, q, 0.13, 0.1)
        for k, v in s.timers().items(): acc[k] = acc.get(k, 0) + v / kicks
    s.enable_timers(False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(kicks): s.kick_device(r, q, 0.13, 0.1)
    e1.record(); torch.cuda.synchronize()
    print(n, nm, {k: round(v * 1e3, 1) for k, v in acc.items()}, "graph us/kick", round(e0.elapsed_time(e1) * 1e3 / kicks, 1), flush=True)
for a in sys.argv[1:]:
    n, nm = a.split(":"); run(int(n), int(nm))
