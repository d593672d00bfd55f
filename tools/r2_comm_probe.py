"""2-GPU timing probe (torchrun): per-kick time of the sharded c2 kick in graph mode with parts of the
communication left out (OCL_SC_DEBUG_SKIP; results are then wrong, only the timing is of interest)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import bench

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
from ocelot_b200.distributed import ShardedSpaceCharge
n, mesh = 1_000_000, 63
p = bench.device_bunch(torch, n, 1234 + rank, dev)
flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device=dev)

def run(label, nvls=True, p2p_rho=False, p2p=True):
    s = ShardedSpaceCharge(step=1, nmesh_xyz=[mesh] * 3, slab=False)
    s.nvls_rho, s.p2p_rho, s.p2p = nvls, p2p_rho, p2p
    s.prepare(None)
    for _ in range(6):
        s.apply(p, 0.1)
    dist.barrier(); torch.cuda.synchronize()
    K = 200
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    for a, b in ev:
        flush.fill_(1.0); a.record(); s.apply(p, 0.1); b.record()
    dist.barrier(); torch.cuda.synchronize()
    ms = sum(a.elapsed_time(b) for a, b in ev) / K
    t = torch.tensor([ms], device=dev, dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    # stage timers, no graph
    s.use_graph = False
    sol = s._engine.solver
    sol.enable_timers(True); acc = {}
    for _ in range(50):
        flush.fill_(1.0); s.apply(p, 0.1)
        for k, v in sol.timers().items(): acc[k] = acc.get(k, 0) + v / 50
    sol.enable_timers(False)
    if rank == 0:
        print(f"{label:28s} graph {t.item()*1e3:7.1f} us/kick   stages(us) " + " ".join(f"{k}={v*1e3:.1f}" for k, v in acc.items()), flush=True)
    s.finalize(); s._engine = None
    import gc; gc.collect(); torch.cuda.synchronize()

mode = os.environ.get("PROBE_MODE", "nvls")
run(f"{mode} skip={os.environ.get('OCL_SC_DEBUG_SKIP','0')}", nvls=(mode == "nvls"), p2p_rho=(mode == "p2p_rho"), p2p=(mode != "nccl_all"))
dist.destroy_process_group()
