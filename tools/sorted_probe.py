"""What would cell-ordered particle storage buy?  Same bunch, caller's (random) order vs sorted by mesh
cell, stage timers of the kick.  (Measurement for DESIGN.md section 8; the product never reorders.)"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ocelot_b200 import native


def stage_times(s, r, q, kicks=20):
    for _ in range(3):
        s.kick_device(r, q, 0.13, 0.1)
    s.enable_timers(True)
    acc = {}
    for _ in range(kicks):
        s.kick_device(r, q, 0.13, 0.1)
        for k, v in s.timers().items():
            acc[k] = acc.get(k, 0) + v / kicks
    s.enable_timers(False)
    return {k: round(v * 1e3, 1) for k, v in acc.items()}


for a in sys.argv[1:]:
    n, nm = (int(v) for v in a.split(":"))
    g = torch.Generator(device="cuda").manual_seed(1)
    r = torch.empty((6, n), dtype=torch.float64, device="cuda")
    sig = [1e-4, 2e-5, 1e-4, 2e-5, 1e-3, 1e-4]
    for k in range(6):
        r[k] = torch.randn(n, generator=g, device="cuda", dtype=torch.float64) * sig[k]
    q = torch.full((n,), 250e-12 / n, dtype=torch.float64, device="cuda")
    s = native.Solver(0, (nm,) * 3)
    print(n, nm, "caller order ", stage_times(s, r, q), flush=True)
    key = torch.zeros(n, dtype=torch.int64, device="cuda")
    for row in (0, 2, 4):                                  # x slowest, z (tau) fastest, like the mesh
        c = r[row]
        idx = ((c - c.min()) / (c.max() - c.min()) * (nm - 3)).floor().long().clamp_(0, nm - 1)
        key = key * nm + idx
    perm = torch.argsort(key)
    rs = r[:, perm].contiguous()
    print(n, nm, "cell order   ", stage_times(s, rs, q), flush=True)
    # coarser locality: sorted by 4x4x4 super-cells only
    perm2 = torch.argsort((key // (nm * nm) // 4) * 10000 + ((key // nm) % nm // 4) * 100 + (key % nm) // 4)
    rs2 = r[:, perm2].contiguous()
    print(n, nm, "4^3 tile order", stage_times(s, rs2, q), flush=True)
