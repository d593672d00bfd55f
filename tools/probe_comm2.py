"""Stage timers of the sharded kick with NCCL vs fused-peer rho (run under torchrun)."""
import faulthandler, os, sys
faulthandler.dump_traceback_later(100, exit=True)
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ocelot_b200 import DeviceParticleArray
from ocelot_b200 import distributed as D
rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 1_000_000
g = torch.Generator(device="cuda").manual_seed(5 + rank)
p = DeviceParticleArray(n)
sig = [1e-4, 2e-5, 1e-4, 2e-5, 1e-3, 1e-4]
for k in range(6):
    p.rparticles[k] = torch.randn(n, generator=g, device="cuda", dtype=torch.float64) * sig[k]
p.q_array.fill_(250e-12 / n / world); p.E = 0.13
for label, p2p, prho in (("NCCL", False, False), ("mailbox+NCCL rho", True, False), ("mailbox+peer rho", True, True)):
    sc = D.ShardedSpaceCharge(nmesh_xyz=[63, 63, 63]); sc.p2p = p2p; sc.p2p_rho = prho; sc.use_graph = False; sc.prepare(None)
    for _ in range(4): sc.apply(p, 0.1)
    s = sc._engine.solver
    s.enable_timers(True); acc = {}
    for _ in range(30):
        sc.apply(p, 0.1)
        for k, v in s.timers().items(): acc[k] = acc.get(k, 0) + v / 30
    s.enable_timers(False)
    if rank == 0: print(f"{label:22s}", {k: round(v * 1e3, 1) for k, v in acc.items()}, flush=True)
    sc.finalize()
torch.cuda.synchronize(); dist.barrier(); dist.destroy_process_group()
