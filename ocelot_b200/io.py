"""Particle IO in the reference's npz layout (ocelot/cpbd/io.py:223-241): keys ``rparticles``,
``q_array``, ``E``, ``s`` -- files written here load with the reference's
``load_particle_array`` and vice versa."""
from __future__ import annotations

import numpy as np

from .particles import DeviceParticleArray, ParticleArray


def save_particle_array2npz(filename, p_array):
    """``save_particle_array2npz`` (io.py:223-226) for host or device-resident particle arrays."""
    r, q = p_array.rparticles, p_array.q_array
    if not isinstance(r, np.ndarray):
        r, q = r.cpu().numpy(), q.cpu().numpy()
    np.savez_compressed(filename, rparticles=r, q_array=q, E=p_array.E, s=p_array.s)


def load_particle_array_from_npz(filename, device=None):
    """``load_particle_array_from_npz`` (io.py:229-241); with ``device`` set the particles go straight
    to HBM as a DeviceParticleArray."""
    with np.load(filename) as data:
        r, q = data["rparticles"], data["q_array"]
        E, s = float(data["E"]), float(data["s"])
    host = ParticleArray(r.shape[1])
    host.rparticles[:], host.q_array[:], host.E, host.s = r, q, E, s
    return host if device is None else DeviceParticleArray.from_host(host, device=device)
