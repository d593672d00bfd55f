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
sc = D.ShardedSpaceCharge(nmesh_xyz=[63, 63, 63]); sc.p2p = True; sc.p2p_rho = False; sc.prepare(None); timeit("mailbox x2 + NCCL rho, graph", sc)
ref = p.rparticles.clone()
sc = D.ShardedSpaceCharge(nmesh_xyz=[63, 63, 63]); sc.p2p = True; sc.p2p_rho = True; sc.prepare(None); timeit("mailbox x2 + fused peer rho, graph", sc)
# correctness of the fused path against NCCL on the same input
a = DeviceParticleArray(n); a.rparticles.copy_(ref); a.q_array.copy_(p.q_array); a.E = 0.13
b = DeviceParticleArray(n); b.rparticles.copy_(ref); b.q_array.copy_(p.q_array); b.E = 0.13
s1 = D.ShardedSpaceCharge(nmesh_xyz=[63, 63, 63]); s1.p2p = False; s1.prepare(None)
s2 = D.ShardedSpaceCharge(nmesh_xyz=[63, 63, 63]); s2.p2p_rho = True; s2.prepare(None)
for _ in range(3): s1.apply(a, 0.1); s2.apply(b, 0.1)
torch.cuda.synchronize()
err = float(((a.rparticles - b.rparticles).abs().amax(dim=1) / a.rparticles.std(dim=1)).max())
if rank == 0: print("fused-peer vs NCCL path, max row deviation / rms:", err, flush=True)
s1.finalize(); s2.finalize()
# timing-only variants (physics wrong): no rho all-reduce / no collectives at all
orig = dist.all_reduce
def no_rho(t, *a, **k):
    if t.numel() > 100: return None
    return orig(t, *a, **k)
dist.all_reduce = no_rho
sc = D.ShardedSpaceCharge(nmesh_xyz=[63, 63, 63]); sc.p2p = True; sc.p2p_rho = False; sc.prepare(None); timeit("mailbox x2, NO rho all-reduce (timing only)", sc)
dist.all_reduce = orig
torch.cuda.synchronize(); dist.barrier()
dist.destroy_process_group()
