"""A few kicks of one workload, for ncu (launch list / full capture)."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ocelot_b200 import native

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
nm = int(sys.argv[2]) if len(sys.argv) > 2 else 63
kicks = int(sys.argv[3]) if len(sys.argv) > 3 else 3
g = torch.Generator(device="cuda").manual_seed(1)
r = torch.empty((6, n), dtype=torch.float64, device="cuda")
sig = [1e-4, 2e-5, 1e-4, 2e-5, 1e-3, 1e-4]
for k in range(6):
    r[k] = torch.randn(n, generator=g, device="cuda", dtype=torch.float64) * sig[k]
r[5] += 0.01 * r[4] / 1e-3
q = torch.full((n,), 250e-12 / n, dtype=torch.float64, device="cuda")
s = native.Solver(0, (nm, nm, nm))
for _ in range(kicks):
    s.kick_device(r, q, 0.13, 0.1)
torch.cuda.synchronize()
print("ok", s.launch_count())
