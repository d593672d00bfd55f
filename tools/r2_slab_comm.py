"""Multi-GPU (torchrun): time the collectives of the slab-decomposed solve one by one (bus bandwidth of the
all-to-alls).  Prints, per collective, the time (max over ranks) and the bus bandwidth per GPU."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import bench

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
from ocelot_b200.distributed import ShardedSpaceCharge

def timed(fn, reps=10):
    for _ in range(3):
        fn()
    dist.barrier(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record(); torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / reps], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())

for mesh, n in ((int(os.environ.get("MESH", "255")), int(os.environ.get("NP", "2000000"))),):
    p = bench.device_bunch(torch, n, 99 + rank, dev)
    s = ShardedSpaceCharge(step=1, nmesh_xyz=[mesh] * 3, slab=True)
    s.prepare(None); s.use_graph = False
    for _ in range(2):
        s.apply(p, 0.1)
    eng = s._engine; b = eng.buffers; sol = eng.solver
    f = (world - 1) / world
    rows = []
    if eng.nvls is not None:
        ms = timed(lambda: sol.nvls_reduce_rho())
        rows.append(("rho reduce-scatter (in-switch multimem kernel + 2 barriers)", ms, b["rho_slab"].numel() * 8 * world * f))
    ms = timed(lambda: dist.all_to_all_single(b["xchg_b"], b["xchg_a"]))
    rows.append(("all-to-all (y -> x layout), NCCL", ms, b["xchg_a"].numel() * 8 * f))
    ms = timed(lambda: dist.all_to_all_single(b["xchg_a"], b["xchg_b"]))
    rows.append(("all-to-all (x -> y layout), NCCL", ms, b["xchg_b"].numel() * 8 * f))
    ms = timed(lambda: dist.all_gather_into_tensor(b["phi"], b["phi_slab"]))
    rows.append(("all-gather of phi, NCCL", ms, b["phi"].numel() * 8 * f))
    ms = timed(lambda: s.apply(p, 0.1))
    if rank == 0:
        print(f"world {world}, mesh {mesh}^3, {n} particles per GPU: whole kick {ms*1e3:.0f} us (no graph)")
        for name, t, bytes_ in rows:
            print(f"  {name:62s} {t*1e3:8.1f} us   {bytes_/1e6:8.1f} MB on the wire per GPU   busbw {bytes_/t/1e6:7.1f} GB/s")
    s.finalize()
dist.destroy_process_group()
