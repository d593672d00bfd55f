"""Is the kick launch-bound?  Host time per call vs device time, and a CUDA-graph replay of the same kick."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ocelot_b200 import native

def bunch(n):
    g = torch.Generator(device="cuda").manual_seed(1)
    r = torch.empty((6, n), dtype=torch.float64, device="cuda")
    sig = [1e-4, 2e-5, 1e-4, 2e-5, 1e-3, 1e-4]
    for k in range(6):
        r[k] = torch.randn(n, generator=g, device="cuda", dtype=torch.float64) * sig[k]
    r[5] += 0.01 * r[4] / 1e-3
    q = torch.full((n,), 250e-12 / n, dtype=torch.float64, device="cuda")
    return r, q

for n, nm in ((200_000, 31), (1_000_000, 63), (12_500_000, 127)):
    r, q = bunch(n)
    s = native.Solver(0, (nm,) * 3)
    for _ in range(5):
        s.kick_device(r, q, 0.13, 0.1)
    torch.cuda.synchronize()
    K = 50
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(K):
        s.kick_device(r, q, 0.13, 0.1)
    t_host = (time.perf_counter() - t0) / K; e1.record(); torch.cuda.synchronize()
    t_dev = e0.elapsed_time(e1) / K
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        s.kick_device(r, q, 0.13, 0.1, stream=st)
        st.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=st):
            s.kick_device(r, q, 0.13, 0.1, stream=torch.cuda.current_stream())
        for _ in range(3):
            g.replay()
        torch.cuda.synchronize()
        e0.record(st)
        for _ in range(K):
            g.replay()
        e1.record(st); torch.cuda.synchronize()
    t_graph = e0.elapsed_time(e1) / K
    print(f"n={n} mesh={nm}: host {t_host*1e6:.1f} us/call, device stream {t_dev*1e3:.1f} us/kick, graph replay {t_graph*1e3:.1f} us/kick")
