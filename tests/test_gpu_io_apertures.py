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
    sc = SpaceCharge(nmesh_xyz=[31, 31, 31])
    sc.prepare(None)
    ref = host.rparticles.copy()
    orc.sc_kick(ref, host.q_array, host.E, 0.1, (31, 31, 31), fft="padded")
    sc.apply(dev, 0.1)               # N changed: the kick graph is re-captured for the new size
    got = dev.to_host().rparticles
    for k in range(6):
        assert np.max(np.abs(got[k] - ref[k])) / np.std(ref[k]) < 1e-10


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
