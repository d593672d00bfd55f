"""Device-side neighbours of the space-charge kick in the tracking loop:

* ``apply_map``     -- first/second-order transfer map on a resident bunch
                       (``TransferMap.mul_p_array``, transformations/transfer_map.py:42-53;
                       ``SecondTM.t_apply``, transformations/second_order.py:31-39)
* ``get_envelope``  -- beam moments / Twiss from particles, default path of the reference's
                       ``get_envelope`` (beam/analysis.py:27-222: no bounds, no dispersion correction)

Both run as CUDA kernels behind the C ABI (``ocl_sc_map_apply``, ``ocl_sc_beam_moments``); the
scalar post-processing of the 18 reduced moments (emittances, beta, alpha) is host arithmetic in
the reference's own expressions.
"""
from __future__ import annotations

import numpy as np

from . import constants as _c

_solvers = {}


def _scratch_solver(device: int):
    """A small native handle used only for its reduction scratch (mesh size irrelevant)."""
    from . import native
    s = _solvers.get(device)
    if s is None:
        s = native.Solver(device, (8, 8, 8))
        _solvers[device] = s
    return s


def apply_map(p_array, R, B=None, T=None, delta_e=0.0, length=0.0):
    """``rparticles <- R r + T:rr + B`` on a DeviceParticleArray, then ``E += delta_e``,
    ``s += length`` as ``Transformation.apply`` does (transformations/transformation.py:123-131)."""
    r = p_array.rparticles
    _scratch_solver(r.device.index or 0).map_apply(r, R, B, T)
    p_array.E += delta_e
    p_array.s += length


def cavity_coefficients(v, phi_deg, freq, E, delta_length, length):
    """Scalars of the RF-cavity body map (specification: ``CavityTM.map4cav``, transformations/cavity.py:29-128),
    derived inside the native library (``ocl_sc_cavity_coefficients``, csrc/sc_abi.cu).  Returns
    ``(mode, coef[7], delta_e)`` with coef = [c1, c2, beta0*k, phi, T566, T556, T555]; mode 2 is the drift-like
    branch for a non-physical final energy."""
    from . import native
    return native.cavity_coefficients(v, phi_deg, freq, E, delta_length, length)


def apply_cavity(p_array, R, B, v, phi_deg, freq, delta_length, length, delta_e=None):
    """Body of an RF cavity on a DeviceParticleArray: ``CavityTM.map_function`` for the MAIN map
    (cavity.py:130-132), then ``E += delta_e``, ``s += delta_length`` (transformation.py:129-131)."""
    mode, coef, de = cavity_coefficients(v, phi_deg, freq, float(p_array.E), delta_length, length)
    r = p_array.rparticles
    _scratch_solver(r.device.index or 0).cavity_apply(r, R, B, coef, mode)
    p_array.E += de if delta_e is None else delta_e
    p_array.s += delta_length if delta_length is not None else length


class Moments:
    """Result of ``get_envelope``: the attributes of the reference's ``Twiss`` that the default
    path fills (beam/analysis.py:81-83, :125-222)."""

    def __init__(self):
        self.E = self.q = self.p = self.s = 0.0
        for k in ("x", "px", "y", "py", "tau", "xx", "xpx", "pxpx", "yy", "ypy", "pypy", "tautau", "pp", "xy",
                  "pxpy", "xpy", "ypx", "emit_x", "emit_y", "emit_xn", "emit_yn", "beta_x", "beta_y", "alpha_x",
                  "alpha_y"):
            setattr(self, k, 0.0)

    def __repr__(self):
        return (f"<Moments E={self.E:.6g} emit_x={self.emit_x:.6g} emit_y={self.emit_y:.6g} "
                f"beta_x={self.beta_x:.6g} beta_y={self.beta_y:.6g} sigma_tau={np.sqrt(self.tautau):.6g}>")


def moments_from_sums(m: dict, E=0.0, q=0.0) -> Moments:
    """Scalar post-processing of the 18 reduced moments (beam/analysis.py:179-220)."""
    t = Moments()
    t.E, t.q = float(E), float(q)
    for k, v in m.items():
        setattr(t, k, float(v))
    t.emit_x = np.sqrt(t.xx * t.pxpx - t.xpx ** 2)
    t.emit_y = np.sqrt(t.yy * t.pypy - t.ypy ** 2)
    relgamma = t.E / _c.m_e_GeV
    relbeta = np.sqrt(1 - relgamma ** -2) if relgamma != 0 else 1.
    t.emit_xn = t.emit_x * relgamma * relbeta
    t.emit_yn = t.emit_y * relgamma * relbeta
    t.beta_x = t.xx / t.emit_x
    t.beta_y = t.yy / t.emit_y
    t.alpha_x = -t.xpx / t.emit_x
    t.alpha_y = -t.ypy / t.emit_y
    return t


class EnvelopeRecorder:
    """Beam moments of many steps of a resident tracking run, kept ON THE DEVICE until ``collect()``.

    The reference computes ``get_envelope`` after every step (track.py:482) but hands the list back only when
    tracking is over (track.py:504); reading 19 doubles back per step would put a host synchronisation into every
    step of an otherwise asynchronous loop (kick graph, map kernel, moment kernels).  ``record`` launches the two
    moment passes into the next slot of a device buffer; ``collect`` does one device->host copy and the scalar
    post-processing (analysis.py:179-220) for all recorded steps."""

    SLOT = 24                       # doubles per record (19 used), keeps slots 64-byte aligned
    CHUNK = 128                     # records per device buffer; buffers are never reallocated while writes are pending

    def __init__(self, device):
        self.device = device
        self._chunks = []
        self._meta = []             # (slot or None, E, s)
        self._slots = 0

    def __len__(self):
        return len(self._meta)

    def record(self, p_array, s=0.0):
        """Queue the moments of ``p_array`` as it is now.  ``s`` is what the record's ``.s`` will report: the
        reference's get_envelope leaves Twiss.s at 0 and track() adds the tracked length (track.py:482-484)."""
        import torch
        r = p_array.rparticles
        E, s = float(p_array.E), float(s)
        if r.shape[1] < 3:                       # analysis.py:86-88: an empty Twiss carrying only the energy
            self._meta.append((None, E, 0.0))
            return
        k = self._slots
        self._slots += 1
        if k // self.CHUNK >= len(self._chunks):
            self._chunks.append(torch.zeros(self.CHUNK, self.SLOT, dtype=torch.float64, device=r.device))
        out = self._chunks[k // self.CHUNK][k % self.CHUNK]
        _scratch_solver(r.device.index or 0).beam_moments_device(r, p_array.q_array, out)
        self._meta.append((k, E, s))

    def collect(self):
        """One synchronising device->host copy per buffer; returns the list of ``Moments`` in recording order."""
        from . import native
        host = [c.cpu().numpy() for c in self._chunks]
        res = []
        for slot, E, s in self._meta:
            if slot is None:
                t = Moments()
                t.E = E
            else:
                row = host[slot // self.CHUNK][slot % self.CHUNK]
                t = moments_from_sums(dict(zip(native.Solver.MOMENT_KEYS, row[:18])), E=E, q=float(row[18]))
                t.s = s
            res.append(t)
        return res


def get_envelope(p_array) -> Moments:
    """Beam moments of a DeviceParticleArray (two streaming passes on the GPU, 19 doubles back)."""
    rec = EnvelopeRecorder(p_array.rparticles.device)
    rec.record(p_array)
    return rec.collect()[0]
