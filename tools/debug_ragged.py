import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ocelot_b200 import native
from oracle import sc_oracle as orc
dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
rng = np.random.RandomState(3)
for n, nmesh, E in ((5, (5, 4, 6), 0.005), (1000, (9, 17, 33), 0.05), (257, (33, 8, 8), 1.0)):
    r0 = np.zeros((6, n))
    r0[0], r0[2], r0[4] = rng.randn(n) * 1e-4, rng.randn(n) * 2e-4, rng.randn(n) * 5e-4
    r0[1], r0[3], r0[5] = rng.randn(n) * 1e-5, rng.randn(n) * 1e-5, rng.randn(n) * 1e-3
    q0 = (0.2 + rng.rand(n)) * 1e-12
    s = native.Solver(0, nmesh)
    r_ref = r0.copy(); taps = {}
    orc.sc_kick(r_ref, q0, E, 0.02, nmesh, fft="padded", taps=taps)
    r = dev(r0)
    Ex = s.field_at_particles(r, dev(q0), E).cpu().numpy()
    geo = s.geometry()
    print(n, nmesh, "steps_rel", np.abs(geo["steps"]/taps["steps"]-1), "xoff", geo["X_off"]-taps["X_off"])
    print("  rho maxdiff/qmin", np.max(np.abs(s.rho()-taps["rho"]))/q0.min(), "phi_rel", np.max(np.abs(s.phi()-taps["phi"]))/np.max(np.abs(taps["phi"])))
    print("  E_rel", [float(np.max(np.abs(Ex[:,c]-taps["Exyz"][:,c]))/np.max(np.abs(taps["Exyz"][:,c]))) for c in range(3)])
    for env in ("cufft",):
        os.environ["OCL_SC_SOLVER"]=env
        s2 = native.Solver(0, nmesh)
        del os.environ["OCL_SC_SOLVER"]
        Ex2 = s2.field_at_particles(dev(r0), dev(q0), E).cpu().numpy()
        print("  cufft-path E_rel", [float(np.max(np.abs(Ex2[:,c]-taps["Exyz"][:,c]))/np.max(np.abs(taps["Exyz"][:,c]))) for c in range(3)],
              "phi_rel", np.max(np.abs(s2.phi()-taps["phi"]))/np.max(np.abs(taps["phi"])))
