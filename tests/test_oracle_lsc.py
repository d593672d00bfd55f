"""The LSC oracle (oracle/lsc_oracle.py) against vectors produced by the unmodified reference
``LSC`` (oracle/make_golden_lsc.py -> tests/golden/lsc_*.npz)."""
import os

import numpy as np
import pytest

from oracle import lsc_oracle as lo

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def kicks():
    return np.load(os.path.join(GOLD, "lsc_kicks.npz"))


def case_kwargs(sc):
    return dict(step_profile=bool(sc[6]), smooth_param=float(sc[7]), bounds=(float(sc[8]), float(sc[9])),
                K_max=float(sc[4]), fill_factor=float(sc[5]))


def test_oracle_reproduces_reference_kicks(kicks):
    for name in kicks["names"]:
        sc = kicks[f"{name}_scalars"]
        r = kicks[f"{name}_r_in"].copy()
        st = lo.lsc_kick(r, kicks[f"{name}_q"], float(sc[0]), float(sc[1]), **case_kwargs(sc))
        # same numpy/scipy calls in the same order: identical grid, profile, wake and kick
        assert np.array_equal(st["x"], kicks[f"{name}_x"]), name
        assert st["sigma"] == sc[3], name
        np.testing.assert_allclose(st["current"] / (st["q"] * lo.C_LIGHT), kicks[f"{name}_bunch"], rtol=1e-14)
        np.testing.assert_allclose(st["W"], kicks[f"{name}_W"], rtol=1e-12, atol=1e-12 * np.abs(kicks[f"{name}_W"]).max())
        d_ref = kicks[f"{name}_delta_out"] - kicks[f"{name}_r_in"][5]
        d = r[5] - kicks[f"{name}_r_in"][5]
        assert np.abs(d - d_ref).max() <= 1e-12 * np.abs(d_ref).max(), name
        for row in range(5):
            assert np.array_equal(r[row], kicks[f"{name}_r_in"][row])


def test_oracle_reproduces_reference_lsc_test_lattice():
    g = np.load(os.path.join(GOLD, "lsc_track.npz"))
    for j, k in enumerate(g["kept"]):
        r = g["r_in"][j].copy()
        lo.lsc_kick(r, g["q"], float(g["E"][k]), float(g["dz"][k]), K_max=float(g["K_max"][k]),
                    fill_factor=float(g["fill"][k]))
        d_ref = g["delta_out"][j] - g["r_in"][j][5]
        assert np.abs((r[5] - g["r_in"][j][5]) - d_ref).max() <= 1e-12 * np.abs(d_ref).max()


def test_grid_definition_edge_cases():
    # no smoothing: 1000 bins across the bunch, no widening (analysis.py:298-315)
    a, ds, nb = lo.current_grid(-1.0, 1.0, 0.0)
    assert (a, nb) == (-1.0, 1001) and ds == pytest.approx(2e-3)
    assert lo.smoothing_taps(0.0, ds) is None
    a, ds, nb = lo.current_grid(-1.0, 1.0, 0.1)
    assert a == pytest.approx(-1.3) and nb == int(np.ceil(2.6 / 0.025)) + 1
    G = lo.smoothing_taps(0.1, ds)
    assert len(G) % 2 == 1 and G.sum() == pytest.approx(1.0)


def test_dz_below_threshold_is_a_no_op():
    r = np.random.default_rng(0).normal(size=(6, 100))
    r0 = r.copy()
    assert lo.lsc_kick(r, np.ones(100), 0.13, 5e-11) is None     # sc.py:566-568
    assert np.array_equal(r, r0)
