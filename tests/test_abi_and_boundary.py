"""CPU-only checks: the C-ABI library loads and exports every symbol the header
declares; the host-side mirror behaves like the reference's plugin (API,
side effects, copy semantics); the product never routes through the oracle."""
import copy
import os
import pickle
import re
import sys

import numpy as np
import pytest

from ocelot_b200 import SpaceCharge, PhysProc, ParticleArray, native, constants

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    lib = native.load()
    declared = native.declared_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), name
    assert set(declared) == set(native._SIGNATURES)
    assert lib.ocl_sc_abi_version() == 1


def test_constants_and_fft_size_without_gpu():
    k = native.constants()                           # ocelot/common/globals.py:13-24
    assert k["m_e_eV"] == constants.m_e_eV == 9.10938215e-31 * 299792458.0 ** 2 / 1.6021766208e-19
    assert k["m_e_GeV"] == constants.m_e_GeV
    assert k["epsilon_0"] == constants.epsilon_0
    assert [native.fft_size(n) for n in (4, 31, 63, 64, 65, 127, 255)] == [8, 64, 128, 128, 256, 256, 512]


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(RuntimeError, match="no CUDA device|no CPU fallback"):
        native.Solver(0, (31, 31, 31))
    sc = SpaceCharge()
    p = ParticleArray(10)
    p.E = 0.1
    with pytest.raises(RuntimeError):
        sc.apply(p, 0.1)
    sc.apply(p, 0)                                   # zstep == 0 returns before touching anything (sc.py:210-212)


def test_product_does_not_import_oracle_or_reference():
    pkg = os.path.join(ROOT, "ocelot_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", text, re.M), f
                assert "/root/reference" not in text, f


def test_plugin_api_matches_reference_class():
    sc = SpaceCharge(step=5, nmesh_xyz=[31, 31, 31], random_mesh=True, ste=1)   # unknown kwargs ignored (sc.py:92-102)
    assert isinstance(sc, PhysProc)
    assert sc.step == 5 and sc.nmesh_xyz == [31, 31, 31] and sc.random_mesh is True
    assert sc.low_order_kick is True and sc.random_seed == 10 and sc.debug is False
    for attr in ("start_elem", "end_elem", "indx0", "indx1", "s_start", "s_stop", "z0", "energy"):
        assert hasattr(sc, attr)
    assert repr(sc) == "<SpaceCharge: step=5, nmesh_xyz=[31, 31, 31], random_mesh=True>"   # sc.py:253-258
    d = SpaceCharge()
    assert d.step == 1 and d.nmesh_xyz == [63, 63, 63] and d.random_mesh is False
    sc.finalize()
    # prepare reseeds the GLOBAL numpy RNG (sc.py:104-107)
    sc.prepare(None)
    a = np.random.uniform()
    np.random.seed(10)
    assert a == np.random.uniform()
    sc.random_seed = None
    np.random.seed(123)
    x = np.random.uniform()
    np.random.seed(123)
    sc.prepare(None)
    assert x == np.random.uniform()


def test_deepcopy_and_pickle_drop_native_handles():
    sc = SpaceCharge(step=2, nmesh_xyz=[15, 15, 15])
    sc._solvers = {("fake",): object()}
    sc.counter = 2                                   # injected by Navigator (navi.py:81)
    c = copy.deepcopy(sc)
    assert c._solvers == {} and c.step == 2 and c.nmesh_xyz == [15, 15, 15] and c.counter == 2
    assert c.nmesh_xyz is not sc.nmesh_xyz
    sc._solvers = {}
    r = pickle.loads(pickle.dumps(sc))
    assert r._solvers == {} and r.step == 2


@pytest.mark.skipif(not os.path.isdir("/root/reference/ocelot"), reason="reference checkout not present")
def test_drop_in_under_reference_navigator_and_track(monkeypatch):
    """Unmodified Navigator/track() drive the class (navi.py:63-98, track.py:470-477);
    the native engine is replaced by an oracle-backed double because this box has no GPU."""
    sys.path.insert(0, "/root/reference")
    import logging
    logging.disable(logging.WARNING)
    from ocelot import MagneticLattice, Navigator, Drift, Quadrupole, Marker, track
    from ocelot.cpbd.beam import ParticleArray as RefParticleArray
    from ocelot.cpbd.sc import SpaceCharge as RefSpaceCharge
    from oracle import sc_oracle as orc

    calls = []

    class OracleSolver:                                # stands in for native.Solver in this CPU-only test
        def __init__(self, device, nmesh):
            self.nmesh = nmesh

        def kick_host(self, r, q, E, dz, draws):
            calls.append(dz)
            kw = {} if draws is None else dict(mesh_scale=draws[0], mesh_shift=draws[1])
            orc.sc_kick(r, q, E, dz, self.nmesh, **kw)

    monkeypatch.setattr(SpaceCharge, "_host_device", lambda self: 0)
    monkeypatch.setattr(native, "Solver", OracleSolver)

    def build(sc_cls):
        m1, m2 = Marker(), Marker()
        cell = (m1, Drift(l=0.25), Quadrupole(l=0.2, k1=2.0), Drift(l=0.25), Quadrupole(l=0.2, k1=-2.0), Drift(l=0.1), m2)
        lat = MagneticLattice(cell)
        navi = Navigator(lat)
        navi.unit_step = 0.1
        sc = sc_cls()
        sc.step = 2
        sc.nmesh_xyz = [15, 15, 15]
        navi.add_physics_proc(sc, m1, m2)
        np.random.seed(5)
        p = RefParticleArray(n=3000)
        p.rparticles[:] = orc.gaussian_bunch(3000, energy=0.02, charge=2e-10)[0]
        p.q_array[:] = 2e-10 / 3000
        p.E = 0.02
        return lat, navi, p

    lat, navi, p = build(SpaceCharge)
    track(lat, p, navi, print_progress=False)
    lat2, navi2, p2 = build(RefSpaceCharge)
    track(lat2, p2, navi2, print_progress=False)
    assert len(calls) >= 4
    assert np.array_equal(p.rparticles, p2.rparticles)   # same arithmetic, same bits
    # reset_position deep-copies the process table and calls prepare again (navi.py:142-156)
    navi.reset_position()
    assert navi.process_table.proc_list[0]._solvers == {}


@pytest.mark.skipif(not os.path.isdir("/root/reference/ocelot"), reason="reference checkout not present")
def test_resident_track_loop_drives_reference_navigator(monkeypatch):
    """ocelot_b200.track.track (the device-resident loop) against the reference's track() on a lattice
    with RF cavities (CavityTM entrance/main/exit), quadrupoles and drifts under SecondTM, with space
    charge.  This box has no GPU, so the three device calls are replaced by the oracle's numpy
    restatements; what is checked is the control flow: which parameters are pulled from each map
    (tm.get_params(E)), the energy / path-length bookkeeping and the process scheduling."""
    sys.path.insert(0, "/root/reference")
    import logging
    logging.disable(logging.WARNING)
    from ocelot import MagneticLattice, Navigator, Drift, Quadrupole, Cavity, Marker, SecondTM, track as ref_track
    from ocelot.cpbd.beam import ParticleArray as RefParticleArray
    from ocelot.cpbd.sc import SpaceCharge as RefSpaceCharge
    from oracle import sc_oracle as orc
    import importlib
    T = importlib.import_module("ocelot_b200.track")   # (the package also exports the function `track`)

    class HostDev:                                       # stands in for DeviceParticleArray
        def __init__(self, p):
            self.rparticles, self.q_array = p.rparticles.copy(), p.q_array.copy()
            self.E, self.s = float(p.E), float(p.s)

        @property
        def n(self):
            return self.rparticles.shape[1]

    def apply_map(p, R, B=None, T_=None, delta_e=0.0, length=0.0):
        r = p.rparticles
        if T_ is None:
            r[:] = np.add(np.dot(R, r), np.zeros(6).reshape(6, 1) if B is None else np.asarray(B).reshape(6, 1))
        else:
            r[:] = np.matmul(R, r) + np.einsum('ijk,j...,k...->i...', T_, r, r)
            r[:] = np.add(r, np.asarray(B).reshape(6, 1))
        p.E += delta_e
        p.s += length

    def apply_cavity(p, R, B, v, phi, freq, dlen, length, delta_e=None):
        de = orc.cavity_map(p.rparticles, R, B, v, phi, freq, p.E, dlen, length)
        p.E += de if delta_e is None else delta_e
        p.s += dlen if dlen is not None else length

    class OracleSolver:
        def __init__(self, device, nmesh):
            self.nmesh = nmesh

        def kick_host(self, r, q, E, dz, draws):
            orc.sc_kick(r, q, E, dz, self.nmesh)

    monkeypatch.setattr(T, "apply_map", apply_map)
    monkeypatch.setattr(T, "apply_cavity", apply_cavity)
    from ocelot_b200.beam import moments_from_sums

    def envelope(p):
        m = orc.beam_moments(p.rparticles)
        return moments_from_sums({k: v for k, v in m.items() if not k.startswith("emit")}, E=p.E)

    class HostRecorder:                      # stands in for beam.EnvelopeRecorder (device moment kernels)
        def __init__(self, device):
            self.items = []

        def record(self, p, s=0.0):
            t = envelope(p)
            t.s = s
            self.items.append(t)

        def collect(self):
            return self.items

    monkeypatch.setattr(T, "EnvelopeRecorder", HostRecorder)
    monkeypatch.setattr(T, "DeviceParticleArray", HostDev)
    monkeypatch.setattr(SpaceCharge, "_host_device", lambda self: 0)
    monkeypatch.setattr(native, "Solver", OracleSolver)

    def build(sc_cls):
        m1, m2 = Marker(), Marker()
        cav = Cavity(l=0.5, v=0.01, freq=1.3e9, phi=10.0)
        cell = (m1, Drift(l=0.2), cav, Drift(l=0.1), Quadrupole(l=0.2, k1=3.0), Drift(l=0.2),
                Quadrupole(l=0.2, k1=-3.0, tilt=0.3), Drift(l=0.1), m2)
        lat = MagneticLattice(cell, method={'global': SecondTM})
        navi = Navigator(lat)
        navi.unit_step = 0.05
        sc = sc_cls()
        sc.step = 2
        sc.nmesh_xyz = [15, 15, 15]
        navi.add_physics_proc(sc, m1, m2)
        np.random.seed(5)
        p = RefParticleArray(n=2000)
        p.rparticles[:] = orc.gaussian_bunch(2000, energy=0.02, charge=2e-10)[0]
        p.q_array[:] = 2e-10 / 2000
        p.E = 0.02
        return lat, navi, p

    lat, navi, p = build(RefSpaceCharge)
    tws_ref, p_ref = ref_track(lat, p, navi, print_progress=False)
    lat2, navi2, p2 = build(SpaceCharge)
    dev = HostDev(p2)
    tws, dev = T.track(lat2, dev, navi2)
    assert len(tws) == len(tws_ref)
    assert abs(dev.E - p_ref.E) < 1e-15 and abs(dev.s - p_ref.s) < 1e-12
    for row in range(6):
        assert np.max(np.abs(dev.rparticles[row] - p_ref.rparticles[row])) <= 1e-13 * np.std(p_ref.rparticles[row])
    assert abs(tws[-1].xx / tws_ref[-1].xx - 1) < 1e-12 and abs(tws[-1].emit_x / tws_ref[-1].emit_x - 1) < 1e-9
    assert abs(tws[-1].s - tws_ref[-1].s) < 1e-12


def test_cavity_coefficients_from_the_library_match_the_oracle_map():
    """ocl_sc_cavity_coefficients (host arithmetic inside the C library, no GPU needed) against the oracle's
    restatement of CavityTM.map4cav (cavity.py:29-128) over accelerating, decelerating, zero-crossing,
    partial-length and non-physical cases: the longitudinal map built from the seven scalars must reproduce
    the oracle's rows 4 and 5."""
    from oracle import sc_oracle as orc
    from ocelot_b200.beam import cavity_coefficients
    rng = np.random.RandomState(2)
    r0 = np.zeros((6, 200))
    r0[4], r0[5] = rng.randn(200) * 1e-3, rng.randn(200) * 1e-3
    cases = [(0.02, 18.0, 1.3e9, 0.005, None, 1.0), (0.02, 0.0, 1.3e9, 0.13, 0.25, 1.0), (0.04, 160.0, 3.9e9, 0.5, None, 0.35),
             (0.01, 90.0, 1.3e9, 0.1, None, 1.0), (0.01, -90.0, 1.3e9, 0.1, 0.5, 1.0), (0.0, 30.0, 1.3e9, 0.2, None, 1.0),
             (0.3, 180.0, 1.3e9, 0.1, None, 1.0), (0.02, 25.0, 1.3e9, 0.0, None, 1.0), (0.02, 45.0, 1.3e9, 0.05, 0.0, 0.0)]
    for v, phi, f, E, dl, L in cases:
        want = r0.copy()
        de_ref = orc.cavity_map(want, np.eye(6), np.zeros(6), v, phi, f, E, dl, L)
        mode, c, de = cavity_coefficients(v, phi, f, E, dl, L)
        assert de == pytest.approx(de_ref, rel=1e-15, abs=1e-300)
        x4, x5 = r0[4], r0[5]
        y5 = x5 * c[0] + c[1] * (np.cos(-x4 * c[2] + c[3]) - np.cos(c[3])) if mode == 1 else x5
        y4 = x4 + c[4] * x5 * x5 + c[5] * x4 * x5 + c[6] * x4 * x4
        assert (mode == 2) == (E + de_ref <= 0)
        assert np.max(np.abs(y5 - want[5])) <= 1e-13 * max(1e-30, np.max(np.abs(want[5]))), (v, phi, E)
        assert np.max(np.abs(y4 - want[4])) <= 1e-13 * np.max(np.abs(want[4])), (v, phi, E)


def test_envelope_recorder_host_logic(monkeypatch):
    """beam.EnvelopeRecorder without a GPU: a stand-in for the native handle writes the oracle's moments into the slot
    it is given; the recorder must hand out one slot per record across buffer chunks, keep the caller's s and the
    bunch's E at recording time, give an empty Twiss-like record for a bunch of fewer than 3 particles
    (analysis.py:86-88) and post-process every slot exactly like get_envelope (analysis.py:179-220)."""
    torch = pytest.importorskip("torch")
    import types
    from oracle import sc_oracle as orc
    from ocelot_b200 import beam, native

    class FakeSolver:
        MOMENT_KEYS = native.Solver.MOMENT_KEYS
        calls = 0

        def beam_moments_device(self, r, q, out):
            m = orc.beam_moments(r.numpy())
            out[:18] = torch.tensor([m[k] for k in self.MOMENT_KEYS], dtype=torch.float64)
            out[18] = float(q.sum())
            FakeSolver.calls += 1

    monkeypatch.setattr(beam, "_scratch_solver", lambda device: FakeSolver())
    monkeypatch.setattr(beam.EnvelopeRecorder, "CHUNK", 4)
    rng = np.random.RandomState(0)
    rec = beam.EnvelopeRecorder(None)
    bunches = []
    for k in range(11):
        n = 2 if k == 5 else 500
        p = types.SimpleNamespace(rparticles=torch.from_numpy(rng.randn(6, n) * 1e-4), q_array=torch.full((n,), 1e-12, dtype=torch.float64),
                                  E=0.1 + 0.01 * k, s=99.0)
        rec.record(p, s=0.5 * k)
        bunches.append(p)
    assert len(rec) == 11 and FakeSolver.calls == 10 and len(rec._chunks) == 3
    got = rec.collect()
    for k, (t, p) in enumerate(zip(got, bunches)):
        assert t.E == p.E
        if k == 5:
            assert t.xx == 0.0 and t.s == 0.0
            continue
        ref = orc.beam_moments(p.rparticles.numpy())
        assert t.s == 0.5 * k and abs(t.q - 500e-12) < 1e-24
        for key in ("x", "xx", "xpx", "pxpx", "tautau", "pp", "xpy"):
            assert getattr(t, key) == ref[key], (k, key)
        assert abs(t.emit_x / ref["emit_x"] - 1) < 1e-12 and t.beta_x == t.xx / t.emit_x
