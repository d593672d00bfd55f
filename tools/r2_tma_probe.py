"""Probe (GPU): are the TMA and cp.async row pipelines bit-identical (reduced scalars, geometry, rho)?"""
import os, sys, subprocess, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1:
    import torch
    from ocelot_b200 import native
    from oracle import sc_oracle as orc
    np.random.seed(5)
    r0, q0, E = orc.gaussian_bunch(1_000_000, energy=0.13, charge=250e-12)
    s = native.Solver(0, (63, 63, 63))
    r, q = torch.from_numpy(r0).cuda(), torch.from_numpy(q0).cuda()
    s.field_at_particles(r, q, E)
    g = s.geometry()
    rho = s.rho()
    out = dict(mom=s.collective_buffer(native.BUF_MOMENTUM).cpu().numpy().tolist(),
               ext=s.collective_buffer(native.BUF_EXTENT).cpu().numpy().tolist(),
               steps=g["steps"].tolist(), xoff=g["X_off"].tolist(), gamma0=float(g["gamma0"]),
               rho_sum=float(rho.sum()), rho_hash=float((rho * np.arange(rho.size).reshape(rho.shape)).sum()))
    np.save(sys.argv[1], rho)
    print(json.dumps(out))
else:
    res = {}
    for t in ("0", "1"):
        o = subprocess.run([sys.executable, __file__, f"/tmp/rho{t}.npy"], env=dict(os.environ, OCL_SC_TMA=t), capture_output=True, text=True)
        res[t] = json.loads(o.stdout.strip().splitlines()[-1])
    for k in res["0"]:
        print(k, res["0"][k] == res["1"][k], res["0"][k] if res["0"][k] != res["1"][k] else "", res["1"][k] if res["0"][k] != res["1"][k] else "")
    a, b = np.load("/tmp/rho0.npy"), np.load("/tmp/rho1.npy")
    d = np.argwhere(a != b)
    print("cells that differ:", len(d), d[:6].tolist(), [(a[tuple(i)] - b[tuple(i)]) / 2.5e-16 for i in d[:6]])
