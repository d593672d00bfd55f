"""CPU tests of the LSC host side: the lattice scan of ``prepare`` and the undulator factor against
values recorded from the reference (tests/golden/lsc_track.npz), the per-kick scalars against the
oracle's grid definition, the special functions of csrc/sc_special.h (compiled for the host)
against scipy, and the plugin protocol (deepcopy / pickle / dz threshold)."""
import copy
import ctypes
import os
import pickle
import subprocess

import numpy as np
import pytest

from oracle import lsc_oracle as lo

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


class Elem:
    def __init__(self, l):
        self.l = l


class Undulator(Elem):          # recognised by class name (ocelot_b200/lsc.py::_is_undulator)
    def __init__(self, l, Kx, Ky):
        Elem.__init__(self, l)
        self.Kx, self.Ky = Kx, Ky


class Lattice:
    def __init__(self, seq):
        self.sequence = seq

    def get_sequence_part(self, a, b):
        i, j = self.sequence.index(a), self.sequence.index(b)
        return self.sequence[i:j + 1]


@pytest.fixture(scope="module")
def track_gold():
    return np.load(os.path.join(GOLD, "lsc_track.npz"))


def _prepared(g):
    from ocelot_b200 import LSC
    seq = [Undulator(l, kx, ky) if u else Elem(l)
           for l, u, kx, ky in zip(g["seq_l"], g["seq_is_undulator"], g["seq_Kx"], g["seq_Ky"])]
    lsc = LSC(step=1)
    lsc.start_elem, lsc.end_elem, lsc.s_start = seq[0], seq[-1], float(g["s_start"])
    lsc.prepare(Lattice(seq))
    return lsc


def test_prepare_reproduces_reference_K_profile(track_gold):
    lsc = _prepared(track_gold)
    assert lsc._is_undul_in_beam_line
    assert np.array_equal(lsc.K_s_func(track_gold["s_samples"]), track_gold["K_samples"])


def test_undulator_factor_reproduces_every_reference_kick(track_gold):
    lsc = _prepared(track_gold)
    for z0, dz, K_ref, f_ref in zip(track_gold["z0"], track_gold["dz"], track_gold["K_max"], track_gold["fill"]):
        lsc.z0 = float(z0)
        K, f = lsc.undulator_factor(float(dz))
        assert K == K_ref and f == f_ref


def test_kick_parameters_match_oracle_grid():
    from ocelot_b200 import LSC
    rng = np.random.default_rng(3)
    tau = rng.normal(2e-4, 1e-3, 5000)
    x, y, q = rng.normal(1e-4, 2e-4, 5000), rng.normal(0, 1e-4, 5000), rng.uniform(1, 2, 5000) * 1e-14
    stats = dict(n=5000.0, mean_tau=np.mean(tau), m2_tau=np.sum((tau - np.mean(tau)) ** 2), min_tau=tau.min(),
                 max_tau=tau.max(), sum_q=np.sum(q), sum_x=x.sum(), sum_y=y.sum())
    for sp in (0.1, 0.03, 0.0):
        lsc = LSC(smooth_param=sp, bounds=[-0.5, 0.2])
        prm = lsc.kick_parameters(stats, 0.13, 0.4)
        a, ds, nb = lo.current_grid(tau.min(), tau.max(), np.std(tau) * sp)
        assert prm["nb"] == nb and prm["a"] == pytest.approx(a, rel=1e-14) and prm["ds"] == pytest.approx(ds, rel=1e-13)
        G = lo.smoothing_taps(np.std(tau) * sp, ds)
        assert prm["K"] == (-1 if G is None else (len(G) - 1) // 2)
        assert prm["slice_min"] == pytest.approx(np.mean(tau) - 0.5 * np.std(tau), rel=1e-13)
        assert prm["und"] == 1.0 and prm["q"] == np.sum(q)
        assert prm["pc_ref"] == np.sqrt(0.13 ** 2 / lo.M_E_GEV ** 2 - 1) * lo.M_E_GEV


def test_protocol_deepcopy_pickle_and_threshold(track_gold):
    from ocelot_b200 import LSC, ParticleArray
    lsc = LSC(step=2, step_profile=True, smooth_param=0.2, bounds=[-1, 1], unknown_kwarg=3)
    assert (lsc.step, lsc.step_profile, lsc.smooth_param, lsc.bounds, lsc.slice) == (2, True, 0.2, [-1, 1], None)
    lsc._solvers = {0: object()}
    for clone in (copy.deepcopy(lsc), pickle.loads(pickle.dumps(_prepared(track_gold)))):
        assert clone._solvers == {}
    p = ParticleArray(10)
    p.rparticles[:] = 1.0
    p.E = 0.13
    LSC().apply(p, 1e-11)                       # dz < 1e-10 returns before any device work (sc.py:566-568)
    assert np.all(p.rparticles == 1.0)


def test_host_utilities_match_oracle():
    from ocelot_b200 import LSC
    lsc = LSC()
    w = np.concatenate([[0.0], np.logspace(8, 14, 200)])
    assert np.allclose(lsc.imp_lsc(254.0, 1e-4, w, 0.3), lo.imp_lsc(254.0, 1e-4, w, 0.3), rtol=1e-15, atol=0)
    assert np.allclose(lsc.imp_step_lsc(254.0, 1e-4, w.copy(), 0.3), lo.imp_step_lsc(254.0, 1e-4, w.copy(), 0.3),
                       rtol=1e-15, atol=0)
    s = np.arange(64) * 1e-6
    f, y = lsc.wake2impedance(s, np.exp(-((s - 3e-5) / 1e-5) ** 2))
    s2, w2 = lsc.impedance2wake(f, y)
    assert np.allclose(s2, s) and np.allclose(w2, np.exp(-((s - 3e-5) / 1e-5) ** 2), atol=1e-12)


def test_special_functions_against_scipy(tmp_path):
    """exp(x) E1(x) and K1(x) of csrc/sc_special.h, compiled for the host with g++."""
    from scipy.special import exp1, k1
    src = tmp_path / "sp.cpp"
    src.write_text('#include "%s"\nextern "C" void ev(const double* x, int n, double* e, double* k) {\n'
                   '  for (int i = 0; i < n; ++i) { e[i] = ocl::exp_e1(x[i]); k[i] = ocl::bessel_k1(x[i]); } }\n'
                   % os.path.join(ROOT, "ocelot_b200", "csrc", "sc_special.h"))
    so = tmp_path / "sp.so"
    subprocess.run(["g++", "-O2", "-shared", "-fPIC", "-ffp-contract=off", "-o", str(so), str(src)], check=True)
    lib = ctypes.CDLL(str(so))
    x = np.concatenate([np.logspace(-16, np.log10(40), 3000), np.linspace(0.9, 1.1, 201), np.logspace(-12, 2.8, 3000)])
    e, k = np.empty_like(x), np.empty_like(x)
    vp = ctypes.c_void_p
    lib.ev(vp(x.ctypes.data), len(x), vp(e.ctypes.data), vp(k.ctypes.data))
    m = x <= 40
    ref = np.exp(x[m]) * exp1(x[m])
    assert np.max(np.abs(e[m] - ref) / ref) < 5e-15
    kk = k1(x)
    ok = kk > 1e-300
    assert np.max(np.abs(k[ok] - kk[ok]) / kk[ok]) < 5e-15


def _emulated_device_kick(r, params):
    """What csrc/sc_lsc.cu does with the host-derived parameters, in numpy (oracle pieces): used
    where no GPU is present to check the host logic end to end."""
    nb, a, ds, K = int(params["nb"]), params["a"], params["ds"], int(params["K"])
    tau = r[4]
    sl = (tau >= params["slice_min"]) & (tau < params["slice_max"])
    if params["step_profile"]:
        sigma = min(np.ptp(r[0][sl]), np.ptp(r[2][sl])) / 2
    else:
        sigma = (np.std(r[0][sl]) + np.std(r[2][sl])) / 2.
    C = lo.cic_counts(tau, a, ds, nb)
    if K >= 0:
        G = np.exp(-0.5 * (np.arange(-K, K + 1) * ds / params["sigma_s"]) ** 2)
        G = G / G.sum()
        C = np.convolve(C, G)[K:nb + K]
    x = np.arange(nb) * ds + a
    bunch = params["v"] * C / (ds * C.sum()) / lo.C_LIGHT
    und = params["und"]
    res = lo.wake_lsc(x, bunch, params["gamma"], sigma, params["dz"], bool(params["step_profile"]),
                      K_max=np.sqrt(2 * (und - 1)), fill_factor=1.0 if und != 1 else 0.0)
    W = -res * params["q"]
    r[5] += np.interp(tau, x, W) * 1e-9 / params["pc_ref"]


@pytest.mark.skipif(not os.path.isdir("/root/reference/ocelot"), reason="reference checkout not present")
def test_drop_in_under_reference_navigator_with_undulators(monkeypatch):
    """The unmodified Navigator/track() drive ocelot_b200.LSC through the reference's own LSC test
    lattice (quadrupoles, drifts, two undulators): lattice scan in prepare(), z0 injection, undulator
    factor and per-kick scalars are the product's host code; the device kernels are emulated with
    numpy because this box has no GPU.  Result against the reference LSC on the same bunch."""
    import sys
    sys.path.insert(0, "/root/reference")
    import logging
    logging.disable(logging.WARNING)
    from ocelot import MagneticLattice, Navigator, Drift, Quadrupole, Marker, Undulator, track
    from ocelot.cpbd.beam import generate_parray, Twiss
    from ocelot.cpbd.sc import LSC as RefLSC
    from ocelot_b200 import LSC

    kicks = []

    def apply_host(self, r, q_array, E, dz):
        tau = r[4]
        stats = dict(n=float(len(tau)), mean_tau=np.mean(tau), m2_tau=np.sum((tau - np.mean(tau)) ** 2),
                     min_tau=tau.min(), max_tau=tau.max(), sum_q=np.sum(q_array), sum_x=r[0].sum(), sum_y=r[2].sum())
        params = self.kick_parameters(stats, E, dz)
        kicks.append(params["und"])
        _emulated_device_kick(r, params)

    monkeypatch.setattr(LSC, "_apply_host", apply_host)

    def run(cls):
        tws0 = Twiss(beta_x=6.6, beta_y=16.4, emit_xn=0.5e-6, emit_yn=0.5e-6, E=1)
        d, qf, qd = Drift(l=1), Quadrupole(l=0.5, k1=0.6), Quadrupole(l=0.25, k1=-0.6)
        u = Undulator(lperiod=0.04, nperiods=50, Kx=4, Ky=0.)
        m1, m2 = Marker(), Marker()
        lat = MagneticLattice((m1, qd, d, u, d, qf, d, u, d, qd, m2))
        np.random.seed(10)
        p = generate_parray(sigma_tau=3e-6, sigma_p=1e-4, chirp=0.01, charge=250e-12, nparticles=4000, tws=tws0)
        navi = Navigator(lat, unit_step=0.1)
        navi.add_physics_proc(cls(step=1), m1, m2)
        track(lat, p, navi, print_progress=False)
        return p.rparticles

    got, ref = run(LSC), run(RefLSC)
    assert len(kicks) > 50 and max(kicks) > 1.0 and min(kicks) == 1.0      # kicks inside and outside undulators
    d_ref = ref[5]
    assert np.abs(got[5] - d_ref).max() <= 1e-10 * np.abs(d_ref).max()
    assert np.abs(got[4] - ref[4]).max() <= 1e-10 * np.abs(ref[4]).max()


@pytest.mark.skipif(not os.path.isdir("/root/reference/ocelot"), reason="reference checkout not present")
def test_install_swaps_both_reference_classes():
    import sys
    sys.path.insert(0, "/root/reference")
    import logging
    logging.disable(logging.WARNING)
    import ocelot
    import ocelot.cpbd.sc as ref_sc
    import ocelot_b200
    import ocelot.utils.section_track as st          # star-imported the classes before install
    saved = (ref_sc.SpaceCharge, ref_sc.LSC)
    assert st.SpaceCharge is saved[0] and st.LSC is saved[1]
    try:
        sc_cls, lsc_cls = ocelot_b200.install()
        assert ref_sc.SpaceCharge is ocelot_b200.SpaceCharge is sc_cls and ocelot.SpaceCharge is sc_cls
        assert ref_sc.LSC is ocelot_b200.LSC is lsc_cls and ocelot.LSC is lsc_cls
        assert st.SpaceCharge is sc_cls and st.LSC is lsc_cls           # section_track.py:365,383 compare these
        assert ref_sc.SpaceCharge().nmesh_xyz == [63, 63, 63] and ref_sc.LSC().smooth_param == 0.1
        ocelot_b200.install()                                           # idempotent
    finally:
        ocelot_b200.uninstall()
    assert (ref_sc.SpaceCharge, ref_sc.LSC) == saved and ocelot.SpaceCharge is saved[0] and ocelot.LSC is saved[1]
    assert st.SpaceCharge is saved[0] and st.LSC is saved[1]
