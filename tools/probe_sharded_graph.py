"""2-rank NCCL probe of the graph-captured sharded kick (run under torchrun with a timeout)."""
import faulthandler, os, sys, time
faulthandler.dump_traceback_later(90, exit=True)
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ocelot_b200 import DeviceParticleArray
from ocelot_b200.distributed import ShardedSpaceCharge

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
ref = DeviceParticleArray(n); ref.rparticles.copy_(p.rparticles); ref.q_array.copy_(p.q_array); ref.E = 0.13
sc = ShardedSpaceCharge(nmesh_xyz=[63, 63, 63]); sc.prepare(None)
sd = ShardedSpaceCharge(nmesh_xyz=[63, 63, 63]); sd.prepare(None); sd.use_graph = False
print(rank, "direct kick", flush=True)
for _ in range(3):
    sd.apply(ref, 0.1)
torch.cuda.synchronize(); print(rank, "graph kick 1 (capture)", flush=True)
sc.apply(p, 0.1)
torch.cuda.synchronize(); print(rank, "graph kick 2 (replay)", flush=True)
sc.apply(p, 0.1); sc.apply(p, 0.1)
torch.cuda.synchronize()
err = float((p.rparticles - ref.rparticles).abs().max() / ref.rparticles.abs().max())
print(rank, "graph vs direct max rel diff", err, flush=True)
for name, obj, q in (("direct", sd, ref), ("graph", sc, p)):
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        obj.apply(q, 0.1)
    e1.record(); torch.cuda.synchronize()
    print(rank, name, "us/kick", e0.elapsed_time(e1) * 1e3 / 50, flush=True)
sc._graph = None
torch.cuda.synchronize(); dist.barrier()
print(rank, "destroying", flush=True)
dist.destroy_process_group()
print(rank, "done", flush=True)
