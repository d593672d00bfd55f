"""A few LSC kicks of one workload, for ncu."""
import os, sys, types
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ocelot_b200 import LSC

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
step_profile = len(sys.argv) > 2 and sys.argv[2] == "step"
g = torch.Generator(device="cuda").manual_seed(1)
r = torch.empty((6, n), dtype=torch.float64, device="cuda")
sig = [1e-4, 2e-5, 1e-4, 2e-5, 1e-3, 1e-4]
for k in range(6):
    r[k] = torch.randn(n, generator=g, device="cuda", dtype=torch.float64) * sig[k]
q = torch.full((n,), 250e-12 / n, dtype=torch.float64, device="cuda")
p = types.SimpleNamespace(rparticles=r, q_array=q, E=0.13)
lsc = LSC(step_profile=step_profile)
for _ in range(3):
    lsc.apply(p, 0.1)
torch.cuda.synchronize()
print("ok", lsc.last_params["nb"])
