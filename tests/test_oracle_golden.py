"""The oracle (oracle/sc_oracle.py) against outputs of the UNMODIFIED reference
stored under tests/golden/ by oracle/make_golden.py.  CPU only."""
import numpy as np
import pytest

from oracle import sc_oracle as orc


def relmax(a, b):
    return float(np.max(np.abs(a - b)) / np.max(np.abs(b)))


def test_constants_bits():
    # ocelot/common/globals.py:19-24
    assert orc.M_E_EV == 9.10938215e-31 * 299792458.0 ** 2 / 1.6021766208e-19
    assert abs(orc.M_E_EV - 510998.8671) < 1e-3
    assert abs(orc.EPS_0 - 8.854187817620e-12) < 1e-23


def test_mad_transforms_bit_exact(golden):
    g = golden("kat_small.npz")
    gamref = float(g["E"]) / orc.M_E_GEV
    xp = orc.mad_to_cartesian(g["r_in"], gamref)
    assert np.array_equal(xp, g["xp"])
    back = orc.cartesian_to_mad(xp, np.zeros_like(xp), gamref)
    assert np.array_equal(back, g["mad_roundtrip"])


def test_igf_kernel_bit_exact(golden):
    g = golden("kat_small.npz")
    K1 = orc.igf_kernel(g["rho"].shape, g["steps"])
    assert np.array_equal(K1, g["K1"])
    p = golden("kat_poisson.npz")
    assert np.array_equal(orc.igf_kernel(p["rho"].shape, p["steps"]), p["K1"])


def test_potential_reference_fft_bit_exact(golden):
    p = golden("kat_poisson.npz")
    phi = orc.poisson_potential(p["rho"], p["steps"], fft="reference")
    assert np.array_equal(phi, p["phi"])
    g = golden("kat_small.npz")
    assert np.array_equal(orc.poisson_potential(g["rho"], g["steps"]), g["phi"])


def test_potential_padded_fft_is_same_convolution(golden):
    for name in ("kat_poisson.npz", "kat_small.npz", "kat_c1_31.npz"):
        g = golden(name)
        phi = orc.poisson_potential(g["rho"], g["steps"], fft="padded", workers=2)
        assert relmax(phi, g["phi"]) < 5e-14, name


def test_stage_taps_small(golden):
    g = golden("kat_small.npz")
    taps = {}
    r = g["r_in"].copy()
    orc.sc_kick(r, g["q"], float(g["E"]), float(g["dz"]), g["nmesh"], taps=taps)
    assert np.array_equal(taps["steps"], g["steps"])
    assert np.array_equal(taps["rho"], g["rho"])
    assert taps["gamma0"] == float(g["gamma0"])
    assert np.array_equal(taps["phi"], g["phi"])
    # own trilinear vs scipy.ndimage.map_coordinates inside the reference
    assert np.array_equal(taps["Exyz"], g["Exyz"])
    assert np.array_equal(r, g["r_out"])


def test_random_mesh_draw_order(golden):
    g = golden("kat_randmesh.npz")
    np.random.seed(10)                      # SpaceCharge.prepare, sc.py:104-107
    scale = np.random.uniform(low=1, high=1.1)      # sc.py:175
    shift = np.random.uniform(low=-0.5, high=0.5)   # sc.py:185
    assert scale == g["draws"][0] and shift == g["draws"][1]
    taps = {}
    r = g["r_in"].copy()
    orc.sc_kick(r, g["q"], float(g["E"]), float(g["dz"]), g["nmesh"], mesh_scale=scale, mesh_shift=shift, taps=taps)
    assert np.array_equal(taps["steps"], g["steps"])
    assert np.array_equal(taps["rho"], g["rho"])
    assert np.array_equal(r, g["r_out"])


def test_kick_c1_31(golden):
    g = golden("kat_c1_31.npz")
    r = g["r_in"].copy()
    taps = {}
    orc.sc_kick(r, g["q"], float(g["E"]), float(g["dz"]), g["nmesh"], taps=taps)
    assert np.array_equal(taps["rho"], g["rho"])
    assert np.array_equal(r, g["r_out"])
    # padded real FFT: same kick to round-off
    r2 = g["r_in"].copy()
    taps2 = {}
    orc.sc_kick(r2, g["q"], float(g["E"]), float(g["dz"]), g["nmesh"], fft="padded", workers=2, taps=taps2)
    for c in range(3):
        assert np.max(np.abs(taps2["Exyz"][:, c] - g["Exyz"][:, c])) / np.max(np.abs(g["Exyz"][:, c])) < 1e-12
    for row in range(6):
        assert np.max(np.abs(r2[row] - g["r_out"][row])) / np.std(g["r_out"][row]) < 1e-12


def test_gaussian_bunch_matches_generate_parray(golden):
    g = golden("kat_c1_31.npz")
    np.random.seed(int(g["seed"]))
    r, q, E = orc.gaussian_bunch(g["r_in"].shape[1], energy=0.13, charge=250e-12)
    assert np.array_equal(r, g["r_in"])
    assert np.array_equal(q, g["q"])
    assert E == float(g["E"])


def test_zero_step_is_noop(golden):
    g = golden("kat_small.npz")
    r = g["r_in"].copy()
    orc.sc_kick(r, g["q"], float(g["E"]), 0.0, g["nmesh"])
    assert np.array_equal(r, g["r_in"])


@pytest.mark.parametrize("which", ["first", "last"])
def test_injector_kicks_63(golden, which):
    """Two kicks lifted from the reference's own golden-test path
    (space_charge_test.py:51-66): 6.5 MeV and ~150 MeV, 63^3 mesh."""
    g = golden("kat_injector_63.npz")
    r = g[f"r_in_{which}"].copy()
    orc.sc_kick(r, g["q"], float(g[f"E_{which}"]), float(g[f"dz_{which}"]), g["nmesh"], fft="padded", workers=4)
    ref = g[f"r_out_{which}"]
    for row in range(6):
        assert np.max(np.abs(r[row] - ref[row])) / np.std(ref[row]) < 1e-12


def test_trilinear_outside_is_zero():
    F = np.arange(27, dtype=float).reshape(3, 3, 3) + 1
    c = np.array([-0.01, 0.0, 2.0, 2.01, 1.5])
    one = np.ones_like(c)
    out = orc.trilinear(F, c, one, one)
    assert out[0] == 0.0 and out[3] == 0.0
    assert out[1] == F[0, 1, 1] and out[2] == F[2, 1, 1]
    assert out[4] == 0.5 * F[1, 1, 1] + 0.5 * F[2, 1, 1]


def _moment_err(got, ref_row, keys):
    """north-star tolerance model: second moments / emittances relative; means relative to sigma."""
    ref = dict(zip(keys, ref_row))
    sig = {"x": ref["xx"], "px": ref["pxpx"], "y": ref["yy"], "py": ref["pypy"], "tau": ref["tautau"], "p": ref["pp"]}
    worst = 0.0
    for k in keys:
        if k in sig:
            e = abs(got[k] - ref[k]) / np.sqrt(sig[k])
        else:
            e = abs(got[k] - ref[k]) / abs(ref[k])
        worst = max(worst, e)
    return worst


def test_track_c1_small_replay(golden):
    """Config-1 tracking at reduced size: 20k particles, 2 FODO cells, 20 kicks."""
    g = golden("track_c1_small.npz")
    keys = [str(k) for k in g["moment_keys"]]
    np.random.seed(int(g["seed"]))
    r, q, E = orc.gaussian_bunch(int(g["n"]), energy=float(g["E"]), charge=float(g["charge"]))
    assert np.array_equal(r[:, :64], g["r0_head"])
    assert _moment_err(orc.beam_moments(r), g["moments"][0], keys) < 1e-13
    worst = [0.0]

    def check(step, rr):
        worst[0] = max(worst[0], _moment_err(orc.beam_moments(rr), g["moments"][step + 1], keys))

    orc.replay_track(r, q, E, g["R"], g["B"], g["map_step"], g["kick_dz"], g["nmesh"],
                     lambda rr, qq, EE, dz, nm: orc.sc_kick(rr, qq, EE, dz, nm, fft="padded", workers=4), check)
    assert worst[0] < 1e-9, worst[0]
    stride = int(g["sample_stride"])
    for row in range(6):
        ref = g["r_final_sample"][row]
        assert np.max(np.abs(r[row, ::stride] - ref)) / np.std(ref) < 1e-10


def test_track_second_order_replay(golden):
    """Same lattice with SecondTM maps (R, T, B recorded from the reference): pins the
    oracle's second-order map restatement (second_order.py:31-39, tm_utils.py:54-55)."""
    g = golden("track_second_order.npz")
    keys = [str(k) for k in g["moment_keys"]]
    np.random.seed(int(g["seed"]))
    r, q, E = orc.gaussian_bunch(int(g["n"]), energy=float(g["E"]), charge=float(g["charge"]))
    worst = [0.0]

    def check(step, rr):
        worst[0] = max(worst[0], _moment_err(orc.beam_moments(rr), g["moments"][step + 1], keys))

    orc.replay_track(r, q, E, g["R"], g["B"], g["map_step"], g["kick_dz"], g["nmesh"],
                     lambda rr, qq, EE, dz, nm: orc.sc_kick(rr, qq, EE, dz, nm, fft="padded", workers=4), check,
                     T=g["T"])
    assert worst[0] < 1e-9, worst[0]
    stride = int(g["sample_stride"])
    for row in range(6):
        ref = g["r_final_sample"][row]
        assert np.max(np.abs(r[row, ::stride] - ref)) / np.std(ref) < 1e-10


def test_moment_postprocessing_matches_oracle():
    """Host arithmetic of ocelot_b200.beam.moments_from_sums (analysis.py:179-220)."""
    from ocelot_b200.beam import moments_from_sums
    rng = np.random.RandomState(0)
    r = rng.randn(6, 5000) * np.array([1e-4, 2e-5, 1e-4, 2e-5, 1e-3, 1e-3]).reshape(6, 1)
    m = orc.beam_moments(r)
    sums = {k: m[k] for k in m if not k.startswith("emit")}
    t = moments_from_sums(sums, E=0.13, q=1e-9)
    assert t.emit_x == m["emit_x"] and t.emit_y == m["emit_y"]
    assert t.beta_x == m["xx"] / m["emit_x"] and t.alpha_y == -m["ypy"] / m["emit_y"]


def test_reference_golden_test_replayed_by_oracle(golden):
    """The reference's own golden test (test_track_with_sp: 16 cavities, quads, 193 SC kicks at 63^3)
    replayed from its recorded maps by the oracle: pins the cavity-map restatement and the
    whole chain against the reference run AND the reference's JSON golden particles."""
    g = golden("track_injector_golden.npz")
    r = g["r0"].copy()
    E = orc.replay_recorded_maps(r, g["q"], float(g["E0"]), g,
                                 lambda rr, qq, EE, dz, nm: orc.sc_kick(rr, qq, EE, dz, nm, fft="padded", workers=4))
    assert abs(E - float(g["E_final"])) < 1e-12
    ref = g["reference_here_final"]
    for row in range(6):
        assert np.max(np.abs(r[row] - ref[row])) / np.std(ref[row]) < 1e-10
    # tolerance of the reference's own test: particles absolute 1e-12 (space_charge_test.py:64)
    assert np.max(np.abs(r - g["json_golden_final"])) < 1e-12
