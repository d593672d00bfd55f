"""The CUDA kick under the UNMODIFIED reference's own Navigator / track().

``oracle/_ref`` holds a byte-for-byte staged copy of the reference package (oracle/build_ref.py;
git-ignored, shipped to the GPU box with the snapshot).  These tests build BASELINE config 1's lattice
(FODO cells, SC kick every 0.1 m) with the reference's own element / lattice / navigator classes and
track the same bunch twice through the reference's ``track()`` (track.py:431-504): once with the
reference's ``SpaceCharge`` (numpy, host), once with ``ocelot_b200.SpaceCharge`` (CUDA through the
C ABI, host-array mode: ``ocl_sc_kick_host``).  Nothing here uses the oracle port.
"""
import copy
import logging
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import build_ref  # noqa: E402

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not build_ref.available(), reason="oracle/_ref not staged (python -m oracle.build_ref)")]


@pytest.fixture(scope="module")
def ref():
    logging.disable(logging.WARNING)
    return build_ref.import_reference()


def _fodo(ref, ncell, k1=5.0):
    seq = [ref.Marker(eid="START")]
    for i in range(ncell):
        seq += [ref.Quadrupole(l=0.2, k1=+k1, eid=f"QF{i}"), ref.Drift(l=0.3, eid=f"DA{i}"),
                ref.Quadrupole(l=0.2, k1=-k1, eid=f"QD{i}"), ref.Drift(l=0.3, eid=f"DB{i}")]
    seq.append(ref.Marker(eid="END"))
    return seq


def _run(ref, sc, n, ncell, nmesh, method=None):
    np.random.seed(1)
    p = ref.generate_parray(nparticles=n, energy=0.13, charge=250e-12)
    lat = ref.MagneticLattice(_fodo(ref, ncell), **({"method": method} if method else {}))
    navi = ref.Navigator(lat)
    navi.unit_step = 0.1
    sc.step = 1
    sc.nmesh_xyz = list(nmesh)
    navi.add_physics_proc(sc, lat.sequence[0], lat.sequence[-1])        # deep-copies the process table
    tws, p = ref.track(lat, p, navi, print_progress=False)
    return tws, p, navi


MOMENTS = ("xx", "xpx", "pxpx", "yy", "ypy", "pypy", "tautau", "pp", "emit_x", "emit_y")


def _compare(tws_a, pa, tws_b, pb):
    assert pa.rparticles.shape == pb.rparticles.shape
    worst = 0.0
    for k in range(6):
        worst = max(worst, float(np.max(np.abs(pa.rparticles[k] - pb.rparticles[k])) / np.std(pb.rparticles[k])))
    mom = 0.0
    assert len(tws_a) == len(tws_b)
    for ta, tb in zip(tws_a, tws_b):
        for k in MOMENTS:
            mom = max(mom, abs(getattr(ta, k) - getattr(tb, k)) / abs(getattr(tb, k)))
        for k, s in (("x", "xx"), ("y", "yy"), ("tau", "tautau"), ("p", "pp")):
            mom = max(mom, abs(getattr(ta, k) - getattr(tb, k)) / np.sqrt(getattr(tb, s)))
    return worst, mom


def test_cuda_class_under_reference_track_config1(ref):
    """Config-1 lattice (10 m FODO, 100 kicks), 50 k particles, 31^3: every step's moments within 1e-9
    and the final particles within 1e-10 of the row rms of the reference's own SpaceCharge."""
    from ocelot_b200 import SpaceCharge
    from ocelot.cpbd.sc import SpaceCharge as RefSpaceCharge
    assert RefSpaceCharge.__module__ == "ocelot.cpbd.sc" and SpaceCharge is not RefSpaceCharge
    tws_g, p_g, navi_g = _run(ref, SpaceCharge(), 50_000, 10, (31, 31, 31))
    tws_r, p_r, _ = _run(ref, RefSpaceCharge(), 50_000, 10, (31, 31, 31))
    assert len(tws_g) == 101
    worst, mom = _compare(tws_g, p_g, tws_r, p_r)
    print(f"track() with the CUDA class vs the reference class: rows {worst:.2e} of rms, moments {mom:.2e}")
    assert worst < 1e-10
    assert mom < 1e-9
    # the process the navigator drove was the CUDA one, and it really created a native handle
    proc = navi_g.process_table.proc_list[0]
    assert type(proc).__module__.startswith("ocelot_b200") and len(proc._solvers) == 1
    navi_g.reset_position()                                               # navi.py:142-156
    assert navi_g.process_table.proc_list[0]._solvers == {}


def test_installed_class_under_second_order_maps(ref):
    """install() swaps the class inside the reference package; a script written against the reference
    (``from ocelot import *`` names) then tracks with the CUDA kick.  SecondTM lattice, random mesh on:
    both classes consume numpy's global RNG identically (sc.py:104-107, :174-175, :184-185)."""
    import ocelot
    import ocelot.cpbd.sc as ref_sc
    import ocelot_b200
    RefSpaceCharge = ref_sc.SpaceCharge
    sc_r = RefSpaceCharge()
    sc_r.random_mesh = True
    tws_r, p_r, _ = _run(ref, sc_r, 20_000, 2, (31, 31, 31), method={"global": ref.SecondTM})
    try:
        ocelot_b200.install()
        assert ocelot.SpaceCharge is ocelot_b200.SpaceCharge and ref_sc.SpaceCharge is ocelot_b200.SpaceCharge
        sc_g = ocelot.SpaceCharge()
        sc_g.random_mesh = True
        tws_g, p_g, _ = _run(ref, sc_g, 20_000, 2, (31, 31, 31), method={"global": ref.SecondTM})
    finally:
        ocelot_b200.uninstall()
    assert ref_sc.SpaceCharge is RefSpaceCharge and ocelot.SpaceCharge is RefSpaceCharge
    worst, mom = _compare(tws_g, p_g, tws_r, p_r)
    print(f"installed class, SecondTM + random mesh: rows {worst:.2e} of rms, moments {mom:.2e}")
    assert worst < 1e-10
    assert mom < 1e-9


def test_deepcopy_scan_reuses_registered_memory_safely(ref):
    """Scans do ``p = deepcopy(p0)`` with one Navigator / SpaceCharge: freshly allocated
    rparticles of the same size, possibly at a recycled address, must be kicked correctly every time."""
    from ocelot_b200 import SpaceCharge
    from ocelot.cpbd.sc import SpaceCharge as RefSpaceCharge
    np.random.seed(3)
    p0 = ref.generate_parray(nparticles=30_000, energy=0.13, charge=250e-12)
    sc_g, sc_r = SpaceCharge(), RefSpaceCharge()
    for sc in (sc_g, sc_r):
        sc.nmesh_xyz = [31, 31, 31]
    want = copy.deepcopy(p0)
    sc_r.apply(want, 0.1)
    for trial in range(6):
        p = copy.deepcopy(p0)
        p.rparticles = p.rparticles * (1.0 + 0.0)          # fresh allocation, same size
        sc_g.apply(p, 0.1)
        for k in range(6):
            assert np.max(np.abs(p.rparticles[k] - want.rparticles[k])) / np.std(want.rparticles[k]) < 1e-10, trial
        del p
