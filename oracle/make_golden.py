#!/usr/bin/env python
"""Generate the golden vectors under tests/golden/ by EXECUTING THE UNMODIFIED
REFERENCE (``/root/reference``, importable only in the build container).

Test infrastructure -- never imported by the product.  Run:

    python oracle/make_golden.py [all|small|c1|injector|track]

Every array written here is an output of the reference's own code
(``ocelot.cpbd.sc.SpaceCharge``, ``ocelot.cpbd.coord_transform``,
``ocelot.cpbd.beam.generate_parray``, ``ocelot.cpbd.track.track``) on seeded
inputs; the oracle (oracle/sc_oracle.py) and the CUDA path are both checked
against them.  The GPU box has no ``/root/reference``; it only sees the .npz
files.
"""
from __future__ import annotations

import copy
import json
import os
import sys

import numpy as np

REF = os.environ.get("OCELOT_REFERENCE", "/root/reference")
if not os.path.isdir(REF):
    sys.exit(f"reference checkout not found at {REF}; golden vectors can only be made in the build container")
sys.path.insert(0, REF)

import logging  # noqa: E402

logging.disable(logging.WARNING)

from ocelot.cpbd.sc import SpaceCharge  # noqa: E402
from ocelot.cpbd.coord_transform import xxstg_2_xp_mad, xp_2_xxstg_mad  # noqa: E402
from ocelot.cpbd.beam import ParticleArray, generate_parray, get_envelope  # noqa: E402
from ocelot.common.globals import m_e_GeV  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
os.makedirs(OUT, exist_ok=True)

MOMENT_KEYS = ("x", "px", "y", "py", "tau", "p", "xx", "xpx", "pxpx", "yy", "ypy", "pypy",
               "tautau", "pp", "xy", "pxpy", "xpy", "ypx", "emit_x", "emit_y")


class TappedSpaceCharge(SpaceCharge):
    """Reference SpaceCharge with its intermediate arrays recorded (no arithmetic changed)."""

    def potential(self, q, steps):
        self.tap_rho = np.array(q, copy=True)
        self.tap_steps = np.array(steps, copy=True)
        phi = SpaceCharge.potential(self, q, steps)
        self.tap_phi = np.array(phi, copy=True)
        return phi

    def el_field(self, X, Q, gamma, nxyz):
        self.tap_xyz = np.array(X, copy=True)
        self.tap_gamma0 = float(gamma)
        E = SpaceCharge.el_field(self, X, Q, gamma, nxyz)
        self.tap_E = np.array(E, copy=True)
        return E


def make_parray(r, q, E):
    p = ParticleArray(n=r.shape[1])
    p.rparticles[:] = r
    p.q_array[:] = q
    p.E = E
    return p


def reference_kick(r, q, E, dz, nmesh, random_mesh=False):
    sc = TappedSpaceCharge()
    sc.nmesh_xyz = list(nmesh)
    sc.random_mesh = random_mesh
    sc.prepare(None)
    p = make_parray(r, q, E)
    sc.apply(p, dz)
    return sc, p.rparticles


def save(name, **arrays):
    path = os.path.join(OUT, name)
    np.savez_compressed(path, **arrays)
    print(f"wrote {path}  ({os.path.getsize(path) / 1e6:.2f} MB)")


# ---------------------------------------------------------------------------
def golden_small():
    """Ragged mesh, non-uniform charges, low energy; every stage tapped."""
    rng = np.random.RandomState(2024)
    n = 4000
    nmesh = (15, 13, 11)
    E = 0.02
    dz = 0.05
    r = np.zeros((6, n))
    r[0] = rng.randn(n) * 2.0e-4
    r[1] = rng.randn(n) * 3.0e-5 + 1.0e-5
    r[2] = rng.randn(n) * 1.5e-4
    r[3] = rng.randn(n) * 2.0e-5 - 2.0e-5
    r[4] = rng.randn(n) * 4.0e-4
    r[5] = rng.randn(n) * 2.0e-3
    q = (0.5 + rng.rand(n)) * 1e-9 / n
    gamref = E / m_e_GeV
    xp = xxstg_2_xp_mad(r, np.zeros((6, n)), gamref)
    back = xp_2_xxstg_mad(xp, np.zeros((6, n)), gamref)
    sc, r_out = reference_kick(r, q, E, dz, nmesh)
    K1 = sc.sym_kernel(sc.tap_rho.shape, sc.tap_steps)
    save("kat_small.npz", r_in=r, q=q, E=E, dz=dz, nmesh=np.array(nmesh), xp=xp, mad_roundtrip=back,
         xyz_rot=sc.tap_xyz, gamma0=sc.tap_gamma0, steps=sc.tap_steps, rho=sc.tap_rho, K1=K1,
         phi=sc.tap_phi, Exyz=sc.tap_E, r_out=r_out)

    # random_mesh: prepare() seeds the global RNG with 10; each kick draws U(1,1.1) then U(-0.5,0.5)
    sc2, r_out2 = reference_kick(r, q, E, dz, nmesh, random_mesh=True)
    np.random.seed(10)
    draws = np.array([np.random.uniform(low=1, high=1.1), np.random.uniform(low=-0.5, high=0.5)])
    save("kat_randmesh.npz", r_in=r, q=q, E=E, dz=dz, nmesh=np.array(nmesh), draws=draws,
         steps=sc2.tap_steps, rho=sc2.tap_rho, Exyz=sc2.tap_E, r_out=r_out2)

    # stand-alone Poisson KAT: random rho on a ragged mesh straight through potential()
    rho = rng.rand(9, 12, 7) * 1e-12
    steps = np.array([1.1e-4, 0.7e-4, 2.3e-3])
    sc3 = SpaceCharge()
    save("kat_poisson.npz", rho=rho, steps=steps, K1=sc3.sym_kernel(rho.shape, steps),
         phi=sc3.potential(rho.copy(), steps))


def golden_c1():
    """Config-1 shape at reduced N: generate_parray Gaussian bunch, 31^3."""
    np.random.seed(1)
    n = 20000
    p = generate_parray(nparticles=n, energy=0.13, charge=250e-12)
    r = p.rparticles.copy()
    q = p.q_array.copy()
    sc, r_out = reference_kick(r, q, p.E, 0.1, (31, 31, 31))
    save("kat_c1_31.npz", seed=1, r_in=r, q=q, E=p.E, dz=0.1, nmesh=np.array((31, 31, 31)),
         gamma0=sc.tap_gamma0, steps=sc.tap_steps, rho=sc.tap_rho, phi=sc.tap_phi, Exyz=sc.tap_E, r_out=r_out)


def golden_injector():
    """Two kicks taken from the reference's own golden test path
    (unit_tests/ebeam_test/space_charge/space_charge_test.py:51-66 with
    space_charge_conf.py): 10k particles, 6.5 MeV -> ~150 MeV, 63^3."""
    import importlib.util
    import pytest  # the conf module uses pytest.fixture

    conf_path = os.path.join(REF, "unit_tests", "ebeam_test", "space_charge", "space_charge_conf.py")
    spec = importlib.util.spec_from_file_location("space_charge_conf", conf_path)
    conf = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(conf)
    from ocelot import MagneticLattice, Navigator, track, SecondTM

    cell = (conf.Marker(), conf.D_14, conf.C_A1_1_1_I1, conf.D_15, conf.C_A1_1_2_I1, conf.D_15, conf.C_A1_1_3_I1,
            conf.D_15, conf.C_A1_1_4_I1, conf.D_15, conf.C_A1_1_5_I1, conf.D_15, conf.C_A1_1_6_I1, conf.D_15,
            conf.C_A1_1_7_I1, conf.D_15, conf.C_A1_1_8_I1, conf.D_22, conf.Q_37_I1, conf.D_23, conf.Q_38_I1)
    lat = MagneticLattice(cell, method={'global': SecondTM})
    p_array = conf.p_array.__wrapped__() if hasattr(conf.p_array, "__wrapped__") else None
    if p_array is None:
        raise RuntimeError("cannot unwrap the p_array fixture")

    record = []

    class Recording(SpaceCharge):
        def apply(self, p_array, zstep):
            before = p_array.rparticles.copy()
            E = float(p_array.E)
            SpaceCharge.apply(self, p_array, zstep)
            record.append((before, E, float(zstep), p_array.rparticles.copy()))

    sc1 = Recording()
    sc1.nmesh_xyz = [63, 63, 63]
    sc1.step = 1
    sc5 = Recording()
    sc5.nmesh_xyz = [63, 63, 63]
    sc5.step = 5
    navi = Navigator(lat)
    navi.add_physics_proc(sc1, lat.sequence[0], conf.C_A1_1_2_I1)
    navi.add_physics_proc(sc5, conf.C_A1_1_2_I1, lat.sequence[-1])
    navi.unit_step = 0.02
    q = p_array.q_array.copy()
    tws, p_end = track(lat, p_array, navi, print_progress=False)
    print("reference golden path: kicks recorded =", len(record))

    # how well does the reference run here agree with its own JSON golden?
    gpath = os.path.join(REF, "unit_tests", "ebeam_test", "space_charge", "ref_results", "test_track_with_sp.json")
    with open(gpath) as f:
        gold = json.load(f)
    gp = gold["p_array"]
    keys = ("x", "px", "y", "py", "tau", "p")
    gold_r = np.array([[prt[k] for prt in gp] for k in keys])
    diffs = {k: float(np.max(np.abs(gold_r[i] - p_end.rparticles[i]))) for i, k in enumerate(keys)}
    rms = {k: float(np.std(gold_r[i])) for i, k in enumerate(keys)}
    print("max |reference-here - JSON golden| per row:", diffs)

    first, last = record[0], record[-1]
    save("kat_injector_63.npz", q=q, nmesh=np.array((63, 63, 63)), n_kicks=len(record),
         r_in_first=first[0], E_first=first[1], dz_first=first[2], r_out_first=first[3],
         r_in_last=last[0], E_last=last[1], dz_last=last[2], r_out_last=last[3],
         json_golden_final=gold_r, reference_here_final=p_end.rparticles)
    with open(os.path.join(OUT, "reference_selfcheck.json"), "w") as f:
        json.dump({"what": "unmodified reference run in the build container vs its own JSON golden "
                           "unit_tests/ebeam_test/space_charge/ref_results/test_track_with_sp.json (final particles)",
                   "kicks": len(record), "max_abs_diff": diffs, "row_std": rms}, f, indent=1)


def golden_injector_track():
    """The reference's own golden test test_track_with_sp (space_charge_test.py:51-66) recorded map by
    map: 8+8 cavities (CavityTM), drifts/quads (SecondTM), two SpaceCharge processes (63^3, step 1 then
    5), unit_step 0.02, 10k particles.  The fixture lets the device path replay the whole run and be
    compared with the reference's JSON golden particles."""
    import importlib.util
    conf_path = os.path.join(REF, "unit_tests", "ebeam_test", "space_charge", "space_charge_conf.py")
    spec = importlib.util.spec_from_file_location("space_charge_conf", conf_path)
    conf = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(conf)
    from ocelot import MagneticLattice, Navigator, SecondTM

    cell = (conf.Marker(), conf.D_14, conf.C_A1_1_1_I1, conf.D_15, conf.C_A1_1_2_I1, conf.D_15, conf.C_A1_1_3_I1,
            conf.D_15, conf.C_A1_1_4_I1, conf.D_15, conf.C_A1_1_5_I1, conf.D_15, conf.C_A1_1_6_I1, conf.D_15,
            conf.C_A1_1_7_I1, conf.D_15, conf.C_A1_1_8_I1, conf.D_22, conf.Q_37_I1, conf.D_23, conf.Q_38_I1)
    lat = MagneticLattice(cell, method={'global': SecondTM})
    p_array = conf.p_array.__wrapped__()
    r0, q0, E0 = p_array.rparticles.copy(), p_array.q_array.copy(), float(p_array.E)
    sc1, sc5 = SpaceCharge(), SpaceCharge()
    sc1.nmesh_xyz = sc5.nmesh_xyz = [63, 63, 63]
    sc1.step, sc5.step = 1, 5
    navi = Navigator(lat)
    navi.add_physics_proc(sc1, lat.sequence[0], conf.C_A1_1_2_I1)
    navi.add_physics_proc(sc5, conf.C_A1_1_2_I1, lat.sequence[-1])
    navi.unit_step = 0.02

    kind, Rs, Bs, Tidx, Ts, cav, dE, dL, map_step, kick_dz, names = [], [], [], [], [], [], [], [], [], [], set()
    step = 0
    for t_maps, dz, proc_list, phys_steps in navi.get_next_step():      # body of track(), track.py:470-477
        for tm in t_maps:
            prm = tm.get_params(p_array.E)
            name = type(tm).__name__
            names.add(name + ":" + tm.tm_type.name)
            dl = tm.delta_length if tm.delta_length is not None else tm.length
            if name == "SecondTM":
                R, T = (prm.R, prm.T) if prm.tilt == 0 else (prm.get_rotated_R(), prm.get_rotated_T())
                kind.append(1); Tidx.append(len(Ts)); Ts.append(np.array(T, dtype=float))
                cav.append([0, 0, 0, np.nan, 0])
            elif name == "CavityTM" and tm.tm_type.name == "MAIN":
                R = prm.get_rotated_R()
                kind.append(2); Tidx.append(-1)
                cav.append([prm.v, prm.phi, prm.freq, np.nan if tm.delta_length is None else tm.delta_length, tm.length])
            else:                                                       # first-order (incl. cavity edges)
                R = prm.get_rotated_R()
                kind.append(0); Tidx.append(-1); cav.append([0, 0, 0, np.nan, 0])
            Rs.append(np.array(R, dtype=float)); Bs.append(np.array(prm.B, dtype=float).reshape(6))
            dE.append(float(tm.get_delta_e())); dL.append(float(dl)); map_step.append(step)
            tm.apply(p_array)
        dzk = 0.0
        for p, z_step in zip(proc_list, phys_steps):
            p.z0 = navi.z0
            p.apply(p_array, z_step)
            dzk = z_step
        assert len(proc_list) <= 1
        kick_dz.append(float(dzk))
        step += 1
    print("golden path maps:", sorted(names), "steps", step, "maps", len(kind), "kicks", int(np.count_nonzero(kick_dz)))
    gpath = os.path.join(REF, "unit_tests", "ebeam_test", "space_charge", "ref_results", "test_track_with_sp.json")
    with open(gpath) as f:
        gp = json.load(f)["p_array"]
    keys = ("x", "px", "y", "py", "tau", "p")
    gold_r = np.array([[prt[k] for prt in gp] for k in keys])
    print("reference-here vs JSON golden, max abs per row:", np.max(np.abs(gold_r - p_array.rparticles), axis=1))
    save("track_injector_golden.npz", r0=r0, q=q0, E0=E0, nmesh=np.array((63, 63, 63)), kind=np.array(kind),
         R=np.array(Rs), B=np.array(Bs), Tidx=np.array(Tidx), T=np.array(Ts), cav=np.array(cav, dtype=float),
         delta_e=np.array(dE), dl=np.array(dL), map_step=np.array(map_step), kick_dz=np.array(kick_dz),
         json_golden_final=gold_r, reference_here_final=p_array.rparticles.copy(), E_final=float(p_array.E))


def _fodo(k1=5.0, ncell=10):
    from ocelot import Quadrupole, Drift, Marker
    seq = [Marker(eid="START")]
    for i in range(ncell):
        seq += [Quadrupole(l=0.2, k1=+k1, eid=f"QF{i}"), Drift(l=0.3, eid=f"DA{i}"),
                Quadrupole(l=0.2, k1=-k1, eid=f"QD{i}"), Drift(l=0.3, eid=f"DB{i}")]
    seq.append(Marker(eid="END"))
    return seq


def _track_fixture(name, n, ncell, nmesh, sample_stride, second_order=False):
    """Config-1 tracking: Gaussian bunch through ncell FODO cells (1 m each),
    first-order maps, SC kick every 0.1 m.  The per-step transfer matrices the
    reference used are stored so the run can be replayed without Ocelot."""
    from ocelot import MagneticLattice, Navigator

    np.random.seed(1)
    p_array = generate_parray(nparticles=n, energy=0.13, charge=250e-12)
    r0 = p_array.rparticles.copy()
    if second_order:
        from ocelot import SecondTM
        lat = MagneticLattice(_fodo(ncell=ncell), method={'global': SecondTM})
    else:
        lat = MagneticLattice(_fodo(ncell=ncell))
    navi = Navigator(lat)
    navi.unit_step = 0.1
    sc = SpaceCharge()
    sc.step = 1
    sc.nmesh_xyz = list(nmesh)
    navi.add_physics_proc(sc, lat.sequence[0], lat.sequence[-1])

    Rs, Bs, Ts, map_step, dzs, moments = [], [], [], [], [], []

    def env(p):
        t = get_envelope(p)
        return [float(getattr(t, k)) for k in MOMENT_KEYS]

    moments.append(env(p_array))
    step = 0
    # the body of track() (track.py:470-485), with the maps recorded
    for t_maps, dz, proc_list, phys_steps in navi.get_next_step():
        for tm in t_maps:
            prm = tm.get_params(p_array.E)
            if second_order:                     # SecondTM.t_apply, second_order.py:31-39 (tilt == 0 here)
                assert prm.tilt == 0
                Rs.append(np.array(prm.R, dtype=float))
                Ts.append(np.array(prm.T, dtype=float))
            else:
                Rs.append(np.array(prm.get_rotated_R(), dtype=float))
            Bs.append(np.array(prm.B, dtype=float).reshape(6))
            map_step.append(step)
            tm.apply(p_array)
        kick_dz = 0.0
        for p, z_step in zip(proc_list, phys_steps):
            p.z0 = navi.z0
            p.apply(p_array, z_step)
            kick_dz = z_step
        dzs.append(kick_dz)
        moments.append(env(p_array))
        step += 1
    print(f"{name}: steps={step} maps={len(Rs)} kicks={int(np.count_nonzero(dzs))}")
    save(name, seed=1, n=n, E=p_array.E, charge=250e-12, nmesh=np.array(nmesh),
         r0_head=r0[:, :64], r0_checksum=np.array([r0.sum(), np.abs(r0).sum(), (r0 * r0).sum()]),
         R=np.array(Rs), B=np.array(Bs), T=np.array(Ts), map_step=np.array(map_step), kick_dz=np.array(dzs),
         moment_keys=np.array(MOMENT_KEYS), moments=np.array(moments),
         sample_stride=sample_stride, r_final_sample=p_array.rparticles[:, ::sample_stride].copy())


def golden_track():
    _track_fixture("track_c1_small.npz", n=20000, ncell=2, nmesh=(31, 31, 31), sample_stride=10)
    _track_fixture("track_c1.npz", n=200000, ncell=10, nmesh=(31, 31, 31), sample_stride=100)
    _track_fixture("track_second_order.npz", n=20000, ncell=2, nmesh=(31, 31, 31), sample_stride=10, second_order=True)


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what in ("all", "small"):
        golden_small()
    if what in ("all", "c1"):
        golden_c1()
    if what in ("all", "injector"):
        golden_injector()
    if what in ("all", "track"):
        golden_track()
    if what in ("all", "injector_track"):
        golden_injector_track()
