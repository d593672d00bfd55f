#!/usr/bin/env python
"""Generate tests/golden/lsc_*.npz by EXECUTING THE UNMODIFIED REFERENCE ``LSC``
(``/root/reference/ocelot/cpbd/sc.py:261-599``; importable only in the build container).

Test infrastructure -- never imported by the product.  Run:

    python oracle/make_golden_lsc.py

Files written:
  lsc_kicks.npz   single ``LSC.apply`` calls on seeded ``generate_parray`` bunches: Gaussian and
                  step-profile impedance, three energies, different smoothing / slice bounds, an
                  undulator factor; stage outputs (current profile, wake, sigma) tapped from the
                  reference's own ``s_to_cur`` / ``wake_lsc`` calls.
  lsc_track.npz   the reference's own LSC test (unit_tests/ebeam_test/long_space_charge):
                  10 k particles tracked through quadrupoles, drifts and two undulators with
                  ``LSC(step=1)``; the bunch entering and leaving three of the kicks, every kick's
                  (z0, dz, K_max, fill_factor), K_s_func samples and the lattice description needed
                  to rebuild ``prepare`` without the reference.
"""
from __future__ import annotations

import os
import sys

import numpy as np

REF = os.environ.get("OCELOT_REFERENCE", "/root/reference")
if not os.path.isdir(REF):
    sys.exit(f"reference checkout not found at {REF}; golden vectors can only be made in the build container")
sys.path.insert(0, REF)

import logging  # noqa: E402

logging.disable(logging.WARNING)

import ocelot.cpbd.sc as ref_sc  # noqa: E402
from ocelot.cpbd.sc import LSC  # noqa: E402
from ocelot.cpbd.beam import ParticleArray, generate_parray, Twiss  # noqa: E402
from ocelot.cpbd.elements import Drift, Quadrupole, Undulator, Marker  # noqa: E402
from ocelot.cpbd.magnetic_lattice import MagneticLattice  # noqa: E402
from ocelot.cpbd.navi import Navigator  # noqa: E402
from ocelot.cpbd.track import track  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


class TappedLSC(LSC):
    """Reference LSC with its intermediate results recorded (no arithmetic changed)."""

    def wake_lsc(self, s, bunch, gamma, sigma, dz, K_max=0, fill_factor=0):
        res = LSC.wake_lsc(self, s, bunch, gamma, sigma, dz, K_max, fill_factor)
        self.tap = dict(x=np.array(s), bunch=np.array(bunch), gamma=float(gamma), sigma=float(sigma),
                        dz=float(dz), K_max=float(K_max), fill_factor=float(fill_factor), res=np.array(res))
        return res


def make_parray(r, q, E):
    p = ParticleArray(n=r.shape[1])
    p.rparticles[:] = r
    p.q_array[:] = q
    p.E = E
    return p


def single_kicks():
    out = {}
    cases = [
        # name, n, energy, sigma_tau, kwargs, dz, (K_max, fill)
        ("gauss_130MeV", 10000, 0.13, 1e-3, {}, 0.5, None),
        ("step_130MeV", 10000, 0.13, 1e-3, {"step_profile": True}, 0.5, None),
        ("gauss_1GeV_short", 10000, 1.0, 3e-6, {"smooth_param": 0.05, "bounds": [-0.6, 0.3]}, 1.0, None),
        ("step_17MeV", 6000, 0.017, 2e-3, {"step_profile": True, "smooth_param": 0.2}, 0.1, None),
        ("gauss_undulator", 6000, 1.0, 3e-6, {}, 0.3, (4.0, 0.37)),
    ]
    names = []
    for seed, (name, n, energy, sig_tau, kw, dz, und) in enumerate(cases, start=3):
        np.random.seed(seed)
        p = generate_parray(sigma_x=1e-4, sigma_px=2e-5, sigma_tau=sig_tau, sigma_p=1e-4, chirp=0.01,
                            charge=250e-12, nparticles=n, energy=energy)
        # off-centre bunch with unequal charges: exercises the shifted sums and sum(q)
        p.rparticles[0] += 3e-4
        p.rparticles[2] -= 1e-4
        p.rparticles[4] += 0.2 * sig_tau
        p.q_array[:] = p.q_array * np.random.uniform(0.5, 1.5, n)
        r_in = p.rparticles.copy()
        lsc = TappedLSC(step=1, **kw)
        if und is not None:
            # the undulator factor enters wake_lsc only through (K_max, fill_factor): inject them the
            # way apply() derives them (sc.py:569-571) with a constant K profile over part of the step
            lsc._is_undul_in_beam_line = True
            frac = und[1]

            def K_s(x, _k=und[0], _f=frac, _dz=dz):
                x = np.asarray(x)
                return np.where(x - (10.0 - _dz) <= _f * _dz, _k, 0.0)
            lsc.K_s_func = K_s
            lsc.z0 = 10.0
        lsc.apply(p, dz)
        t = lsc.tap
        names.append(name)
        out[name + "_r_in"] = r_in
        out[name + "_q"] = p.q_array.copy()
        out[name + "_delta_out"] = p.rparticles[5].copy()
        out[name + "_scalars"] = np.array([energy, dz, t["gamma"], t["sigma"], t["K_max"], t["fill_factor"],
                                           float(kw.get("step_profile", False)), kw.get("smooth_param", 0.1),
                                           kw.get("bounds", [-0.4, 0.4])[0], kw.get("bounds", [-0.4, 0.4])[1]])
        out[name + "_x"] = t["x"]
        out[name + "_bunch"] = t["bunch"]
        out[name + "_W"] = -t["res"] * np.sum(p.q_array)
        for rr in range(6):
            if rr != 5:
                assert np.array_equal(p.rparticles[rr], r_in[rr])      # LSC only touches delta
    out["names"] = np.array(names)
    path = os.path.join(OUT, "lsc_kicks.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path} ({os.path.getsize(path) / 1e6:.2f} MB)")


class RecordingLSC(TappedLSC):
    log = None
    keep = ()

    def apply(self, p_array, dz):
        k = len(self.log["z0"])
        r_in = p_array.rparticles.copy()
        self.tap = None
        LSC.apply(self, p_array, dz)
        self.log["z0"].append(float(self.z0))
        self.log["dz"].append(float(dz))
        self.log["E"].append(float(p_array.E))
        self.log["K_max"].append(self.tap["K_max"] if self.tap else 0.0)
        self.log["fill"].append(self.tap["fill_factor"] if self.tap else 0.0)
        if k in self.keep:
            self.log["r_in"].append(r_in)
            self.log["delta_out"].append(p_array.rparticles[5].copy())


def reference_lsc_test():
    """unit_tests/ebeam_test/long_space_charge/long_space_charge_conf.py lattice and bunch."""
    tws0 = Twiss(beta_x=6.6, beta_y=16.4, emit_xn=0.5e-6, emit_yn=0.5e-6, E=1)
    d = Drift(l=1)
    qf = Quadrupole(l=0.5, k1=0.6)
    qd = Quadrupole(l=0.25, k1=-0.6)
    u = Undulator(lperiod=0.04, nperiods=50, Kx=4, Ky=0.)
    m1, m2 = Marker(), Marker()
    lat = MagneticLattice((m1, qd, d, u, d, qf, d, u, d, qd, m2))
    np.random.seed(10)
    p = generate_parray(sigma_tau=3e-6, sigma_p=1e-4, chirp=0.01, charge=250e-12, nparticles=10000,
                        tws=tws0, shape="gauss")
    lsc = RecordingLSC(step=1)
    navi = Navigator(lat, unit_step=0.1)
    navi.add_physics_proc(lsc, m1, m2)
    proc = navi.process_table.proc_list[0] if hasattr(navi, "process_table") else lsc
    proc.log = dict(z0=[], dz=[], E=[], K_max=[], fill=[], r_in=[], delta_out=[])
    proc.keep = (0, 14, 40)
    q = p.q_array.copy()
    track(lat, p, navi, print_progress=False)
    log = proc.log
    s = np.linspace(proc.s_start, proc.s_stop, num=57)
    seq = lat.get_sequence_part(proc.start_elem, proc.end_elem)
    path = os.path.join(OUT, "lsc_track.npz")
    np.savez_compressed(
        path, z0=np.array(log["z0"]), dz=np.array(log["dz"]), E=np.array(log["E"]), K_max=np.array(log["K_max"]),
        fill=np.array(log["fill"]), kept=np.array(proc.keep), r_in=np.array(log["r_in"]),
        delta_out=np.array(log["delta_out"]), q=q, s_samples=s, K_samples=proc.K_s_func(s),
        s_start=float(proc.s_start),
        seq_l=np.array([e.l for e in seq]),
        seq_is_undulator=np.array([isinstance(e, Undulator) for e in seq]),
        seq_Kx=np.array([getattr(e, "Kx", 0.0) for e in seq]), seq_Ky=np.array([getattr(e, "Ky", 0.0) for e in seq]))
    print(f"wrote {path} ({os.path.getsize(path) / 1e6:.2f} MB); {len(log['z0'])} kicks, "
          f"K_max in {sorted(set(log['K_max']))}")


if __name__ == "__main__":
    single_kicks()
    reference_lsc_test()
