"""Graph-mode kick time only (A/B of launch mechanics)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ocelot_b200 import native
for a in sys.argv[1:]:
    n, nm = (int(v) for v in a.split(":"))
    g = torch.Generator(device="cuda").manual_seed(1)
    r = torch.empty((6, n), dtype=torch.float64, device="cuda")
    sig = [1e-4, 2e-5, 1e-4, 2e-5, 1e-3, 1e-4]
    for k in range(6):
        r[k] = torch.randn(n, generator=g, device="cuda", dtype=torch.float64) * sig[k]
    q = torch.full((n,), 250e-12 / n, dtype=torch.float64, device="cuda")
    s = native.Solver(0, (nm,) * 3)
    for _ in range(5): s.kick_device(r, q, 0.13, 0.1)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for rep in range(3):
        e0.record()
        for _ in range(30): s.kick_device(r, q, 0.13, 0.1)
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e3 / 30)
    print(n, nm, "graph us/kick", round(best, 1), "graph" if os.environ.get("OCL_SC_GRAPH", "1") != "0" else "no-graph", flush=True)
