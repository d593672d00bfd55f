"""Print parity errors and per-stage timings of the CUDA path (run under gpurun)."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ocelot_b200 import native  # noqa: E402
from oracle import sc_oracle as orc  # noqa: E402


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def report(n, nmesh, seed=5, kicks=20):
    np.random.seed(seed)
    r0, q0, E = orc.gaussian_bunch(n, energy=0.13, charge=250e-12)
    s = native.Solver(0, nmesh)
    r, q = dev(r0), dev(q0)
    out = {"n": n, "nmesh": list(nmesh)}
    if n <= 2_000_000:
        taps = {}
        r_ref = r0.copy()
        t0 = time.perf_counter()
        orc.sc_kick(r_ref, q0, E, 0.1, nmesh, fft="padded", workers=8, taps=taps)
        out["oracle_s"] = time.perf_counter() - t0
        Ex = s.field_at_particles(r, q, E).cpu().numpy()
        geo = s.geometry()
        out["steps_rel"] = float(np.max(np.abs(geo["steps"] / taps["steps"] - 1)))
        out["xoff_abs"] = float(np.max(np.abs(geo["X_off"] - taps["X_off"])))
        out["gamma0_rel"] = float(abs(geo["gamma0"] / taps["gamma0"] - 1))
        rho = s.rho()
        out["rho_maxabs_over_q"] = float(np.max(np.abs(rho - taps["rho"])) / q0[0])
        out["phi_rel"] = float(np.max(np.abs(s.phi() - taps["phi"])) / np.max(np.abs(taps["phi"])))
        out["E_rel"] = [float(np.max(np.abs(Ex[:, c] - taps["Exyz"][:, c])) / np.max(np.abs(taps["Exyz"][:, c])))
                        for c in range(3)]
        s.kick_device(r, q, E, 0.1)
        got = r.cpu().numpy()
        out["row_err"] = [float(np.max(np.abs(got[k] - r_ref[k])) / np.std(r_ref[k])) for k in range(6)]
    s.enable_timers(True)
    for _ in range(3):
        s.kick_device(r, q, E, 0.1)
    torch.cuda.synchronize()
    acc = {}
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(kicks):
        s.kick_device(r, q, E, 0.1)
        for k, v in s.timers().items():
            acc[k] = acc.get(k, 0.0) + v / kicks
    out["stage_ms"] = {k: round(v, 4) for k, v in acc.items()}
    s.enable_timers(False)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(kicks):
        s.kick_device(r, q, E, 0.1)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / kicks
    out["ms_per_kick"] = ms
    out["kicks_per_s"] = n / ms * 1e3
    print(json.dumps(out))


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0))
    report(200_000, (31, 31, 31))
    report(1_000_000, (63, 63, 63))
    report(12_500_000, (127, 127, 127), kicks=5)
