"""Device-resident tracking loop: the body of the reference's ``track()``
(ocelot/cpbd/track.py:466-499) with the particles kept in HBM.

``track`` drives an unmodified Ocelot ``Navigator``: ``navi.get_next_step()`` decides the step,
the transfer maps and the physics processes exactly as in the reference; first- and
second-order maps are applied on the device from ``tm.get_params(E)``; any other map or physics
process that needs host arrays falls back to a device->host->device round trip (correct, slow,
logged).  ``replay_track`` runs the same loop from recorded matrices, for machines without Ocelot.
"""
from __future__ import annotations

import logging

import numpy as np

from .beam import apply_map, apply_cavity, EnvelopeRecorder
from .particles import DeviceParticleArray, ParticleArray

logger = logging.getLogger(__name__)

_DEVICE_MAPS = {"TransferMap", "SecondTM", "CavityTM"}


def _apply_tm(tm, p_dev):
    name = type(tm).__name__
    if name in _DEVICE_MAPS and hasattr(tm, "get_params"):
        prm = tm.get_params(p_dev.E)
        dl = tm.delta_length if getattr(tm, "delta_length", None) is not None else tm.length
        if name == "CavityTM" and getattr(tm.tm_type, "name", "") == "MAIN":                  # cavity.py:130-132
            apply_cavity(p_dev, prm.get_rotated_R(), prm.B, prm.v, prm.phi, prm.freq, tm.delta_length, tm.length,
                         delta_e=tm.get_delta_e())
        elif name in ("TransferMap", "CavityTM"):
            apply_map(p_dev, prm.get_rotated_R(), prm.B, None, tm.get_delta_e(), dl)      # transfer_map.py:50-52
        else:
            R, T = prm.R, prm.T                                                           # second_order.py:32-36
            if prm.tilt != 0:
                R, T = prm.get_rotated_R(), prm.get_rotated_T()
            apply_map(p_dev, R, prm.B, T, tm.get_delta_e(), dl)
        return
    # anything else (cavities, kicks, Runge-Kutta ...) needs the reference's own host code
    logger.info("map %s is not device-resident: round trip through the host", name)
    _host_round_trip(p_dev, tm.apply)


def _host_round_trip(p_dev, fn):
    """Run ``fn(host_particle_array)`` on a host copy (an Ocelot ParticleArray when Ocelot is importable:
    Transformation.apply checks the exact class, transformation.py:128)."""
    try:
        from ocelot.cpbd.beam import ParticleArray as RefPA
        host = RefPA(n=p_dev.n)
    except Exception:  # noqa: BLE001
        host = ParticleArray(p_dev.n)
    p_dev.to_host(host)
    fn(host)
    if host.rparticles.shape[1] != p_dev.n:
        raise NotImplementedError("particle loss on the host path changes N; re-create the DeviceParticleArray")
    import torch
    p_dev.rparticles.copy_(torch.from_numpy(np.ascontiguousarray(host.rparticles)))
    p_dev.E, p_dev.s = float(host.E), float(host.s)


def track(lattice, p_array, navi, calc_tws=True, print_progress=False):
    """``track(lattice, p_array, navi)`` of the reference with ``p_array`` a DeviceParticleArray.
    Returns ``(tws_track, p_array)`` like track.py:431-504."""
    if not isinstance(p_array, DeviceParticleArray):
        raise TypeError("p_array must be an ocelot_b200.DeviceParticleArray (use DeviceParticleArray.from_host)")
    # the moments of every step stay on the device until the loop is over: no host synchronisation per step
    rec = EnvelopeRecorder(getattr(p_array.rparticles, "device", None)) if calc_tws else None
    if calc_tws:
        rec.record(p_array)
    L = 0.0
    for t_maps, dz, proc_list, phys_steps in navi.get_next_step():                  # track.py:470
        for tm in t_maps:
            _apply_tm(tm, p_array)                                                   # track.py:471-472
        for p, z_step in zip(proc_list, phys_steps):                                 # track.py:475-477
            p.z0 = navi.z0
            if getattr(p, "device_resident", False) or type(p).__module__.startswith("ocelot_b200"):
                p.apply(p_array, z_step)
            else:
                _host_round_trip(p_array, lambda host, p=p, z=z_step: p.apply(host, z))
        if p_array.n == 0:
            return (rec.collect() if calc_tws else []), p_array
        L += dz
        if calc_tws:
            rec.record(p_array, s=L)                                                 # track.py:482-484
        if print_progress:
            print(f"\rz = {navi.z0} / {lattice.totalLen}", end="")
    for p in navi.get_phys_procs():                                                  # track.py:498-499
        p.finalize()
    return (rec.collect() if calc_tws else []), p_array


def replay_track(p_array, R, B, map_step, kick_dz, sc, T=None, after_step=None):
    """The same loop from recorded maps: for step s apply maps ``m`` with ``map_step[m] == s``
    (``R[m]`` (6,6), ``B[m]`` (6,), optional ``T[m]`` (6,6,6)), then ``sc.apply(p_array, kick_dz[s])``."""
    map_step = np.asarray(map_step)
    for step, dz in enumerate(kick_dz):
        for m in np.nonzero(map_step == step)[0]:
            apply_map(p_array, R[m], B[m], None if T is None else T[m])
        sc.apply(p_array, float(dz))
        if after_step is not None:
            after_step(step, p_array)


def replay_recorded_maps(p_array, g, sc_for_step):
    """Replay a recorded run (fixture layout of oracle/make_golden.py::golden_injector_track) entirely
    on the device: ``g`` holds per map ``kind`` (0 first order, 1 second order, 2 RF cavity body),
    ``R``, ``B``, ``Tidx``/``T``, ``cav`` = (v, phi, freq, delta_length|nan, length), ``delta_e``, ``dl``,
    ``map_step`` and per step ``kick_dz``.  ``sc_for_step(step)`` returns the physics process to
    apply after the maps of that step (or None)."""
    map_step = np.asarray(g["map_step"])
    order = np.argsort(map_step, kind="stable")
    pos = 0
    for step, dz in enumerate(g["kick_dz"]):
        while pos < len(order) and map_step[order[pos]] == step:
            m = int(order[pos])
            pos += 1
            kind = int(g["kind"][m])
            if kind == 2:
                v, phi, freq, dlen, length = (float(x) for x in g["cav"][m])
                apply_cavity(p_array, g["R"][m], g["B"][m], v, phi, freq, None if np.isnan(dlen) else dlen, length,
                             delta_e=float(g["delta_e"][m]))
            else:
                T = g["T"][int(g["Tidx"][m])] if kind == 1 else None
                apply_map(p_array, g["R"][m], g["B"][m], T, float(g["delta_e"][m]), float(g["dl"][m]))
        sc = sc_for_step(step)
        if sc is not None and dz != 0:
            sc.apply(p_array, float(dz))
