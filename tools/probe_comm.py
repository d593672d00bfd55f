"""Decompose the multi-GPU overhead of the sharded kick (run under torchrun)."""
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

def timeit(label, sc):
    for _ in range(4): sc.apply(p, 0.1)
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(100): sc.apply(p, 0.1)
    e1.record(); torch.cuda.synchronize()
    if rank == 0: print(f"{label:40s} {e0.elapsed_time(e1) * 10:.1f} us/kick  mailbox={sc._engine.mailbox is not None}", flush=True)
    sc.finalize()

sc = D.ShardedSpaceCharge(nmesh_xyz=[63, 63, 63]); sc.p2p = False; sc.prepare(None); timeit("NCCL x3, graph", sc)
sc = D.ShardedSpaceCharge(nmesh_xyz=[63, 63, 63]); sc.p2p = True; sc.prepare(None); timeit("mailbox x2 + NCCL rho, graph", sc)
# timing-only variants (physics wrong): no rho all-reduce / no collectives at all
orig = dist.all_reduce
def no_rho(t, *a, **k):
    if t.numel() > 100: return None
    return orig(t, *a, **k)
dist.all_reduce = no_rho
sc = D.ShardedSpaceCharge(nmesh_xyz=[63, 63, 63]); sc.p2p = True; sc.prepare(None); timeit("mailbox x2, NO rho all-reduce (timing only)", sc)
dist.all_reduce = orig
torch.cuda.synchronize(); dist.barrier()
dist.destroy_process_group()
