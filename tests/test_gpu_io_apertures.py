"""Row f4: npz IO in the reference layout and the N-changing apertures on a
device-resident bunch, followed by a kick on the shrunken array."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from oracle import sc_oracle as orc  # noqa: E402


def _bunch(n, seed):
    from ocelot_b200 import ParticleArray, DeviceParticleArray
    np.random.seed(seed)
    r, q, E = orc.gaussian_bunch(n, energy=0.1, charge=1e-10)
    host = ParticleArray(n)
    host.rparticles[:], host.q_array[:], host.E = r, q, E
    return host, DeviceParticleArray.from_host(host)


def test_apertures_match_reference_rule_and_kick_still_works():
    from ocelot_b200 import RectAperture, EllipticalAperture, SpaceCharge
    host, dev = _bunch(60_000, 12)
    for ap in (RectAperture(xmin=-1.5e-4, xmax=2e-4, ymax=1e-4), EllipticalAperture(xmax=1.8e-4, ymax=0.9e-4, dx=1e-5)):
        ap.apply(host, 0.0)          # numpy branch = the reference's own statements
        ap.apply(dev, 0.0)           # device branch
        assert dev.n == host.rparticles.shape[1] and dev.n < 60_000
        assert np.array_equal(dev.to_host().rparticles, host.rparticles)
        assert np.array_equal(dev.q_array.cpu().numpy(), host.q_array)
    assert len(dev.lost_particles) == 60_000 - dev.n
    # copying back into a host container that still has the original size replaces its arrays (the INTEGRATION.md flow)
    full = _bunch(60_000, 12)[0]
    dev.to_host(full)
    assert full.rparticles.shape == (6, dev.n) and np.array_equal(full.rparticles, host.rparticles)
    assert np.array_equal(full.q_array, host.q_array)
    # lost-particle recorder: original indices, in the order the reference deletes them (x plane, y plane, ellipse)
    r0 = _bunch(60_000, 12)[0].rparticles
    ids = np.arange(60_000)
    want = []
    for lost in (lambda r: (r[0] < -1.5e-4) | (r[0] > 2e-4), lambda r: (r[2] < -np.inf) | (r[2] > 1e-4),
                 lambda r: (r[0] - 1e-5) ** 2 / 1.8e-4 ** 2 + (r[2] - 0.0) ** 2 / 0.9e-4 ** 2 > 1.0):
        m = lost(r0[:, ids])
        want += ids[m].tolist()
        ids = ids[~m]
    assert dev.lost_particles == want
    assert [n for _, n in dev.lp_to_pos_hist] and sum(n for _, n in dev.lp_to_pos_hist) == len(want)
    sc = SpaceCharge(nmesh_xyz=[31, 31, 31])
    sc.prepare(None)
    ref = host.rparticles.copy()
    orc.sc_kick(ref, host.q_array, host.E, 0.1, (31, 31, 31), fft="padded")
    sc.apply(dev, 0.1)               # N changed: the kick graph is re-captured for the new size
    got = dev.to_host().rparticles
    for k in range(6):
        assert np.max(np.abs(got[k] - ref[k])) / np.std(ref[k]) < 1e-10


def test_cut_many_tiles_ragged_and_empty():
    """The compaction kernels on sizes around the 1024-particle tile and the 1024-tile scan batch, cuts that keep
    everything / nothing, and a one-particle bunch."""
    from ocelot_b200 import DeviceParticleArray, ParticleArray
    rng = np.random.RandomState(4)
    for n in (1, 5, 1023, 1024, 1025, 200_000, 1024 * 1024 + 77):
        host = ParticleArray(n)
        host.rparticles[:] = rng.randn(6, n)
        host.q_array[:] = rng.rand(n)
        for lo, hi in ((-0.3, 0.8), (-np.inf, np.inf), (5.0, 6.0)):
            dev = DeviceParticleArray.from_host(host)
            lost = dev.cut(0, 4, (lo, hi, 0.0, 0.0))
            keep = ~((host.rparticles[4] < lo) | (host.rparticles[4] > hi))
            assert dev.n == int(keep.sum()) and lost == n - dev.n
            assert np.array_equal(dev.rparticles.cpu().numpy(), host.rparticles[:, keep])
            assert np.array_equal(dev.q_array.cpu().numpy(), host.q_array[keep])
            assert dev.lost_particles == np.nonzero(~keep)[0].tolist()
            assert np.array_equal(dev._current_particle.cpu().numpy(), np.nonzero(keep)[0])


def test_npz_round_trip_in_reference_layout(tmp_path):
    from ocelot_b200 import save_particle_array2npz, load_particle_array_from_npz
    host, dev = _bunch(5_000, 13)
    dev.s = 1.25
    f = str(tmp_path / "beam.npz")
    save_particle_array2npz(f, dev)
    with np.load(f) as z:             # exactly the reference's keys (io.py:223-226)
        assert sorted(z.files) == ["E", "q_array", "rparticles", "s"]
        assert np.array_equal(z["rparticles"], host.rparticles)
    back = load_particle_array_from_npz(f, device="cuda:0")
    assert back.n == 5_000 and back.s == 1.25 and back.E == host.E
    assert torch.equal(back.rparticles, dev.rparticles)
