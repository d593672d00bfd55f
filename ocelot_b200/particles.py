"""Particle containers.

``ParticleArray`` mirrors the data contract of the reference's
``ocelot.cpbd.beam.particle.ParticleArray`` (particle.py:79-84): ``rparticles``
(6, n) fp64 C-order = rows [x, x', y, y', tau, delta], ``q_array`` (n,), ``E``
[GeV], ``s`` [m].  It exists so the package runs where Ocelot is not installed;
``SpaceCharge.apply`` accepts the reference's own class just the same (duck
typing on those four attributes).

``DeviceParticleArray`` keeps ``rparticles`` and ``q_array`` resident in HBM as
torch CUDA tensors between kicks (rows padded to a 128-byte multiple so every
row start is aligned for vector loads).
"""
from __future__ import annotations

import numpy as np


class ParticleArray:
    """Host container, same fields as particle.py:79-84."""

    def __init__(self, n=0):
        self.rparticles = np.zeros((6, n))
        self.q_array = np.zeros(n)
        self.s = 0.0
        self.E = 0.0

    @property
    def n(self):
        return self.rparticles.shape[1]

    def size(self):
        return self.rparticles.shape[1]

    def x(self):
        return self.rparticles[0]

    def px(self):
        return self.rparticles[1]

    def y(self):
        return self.rparticles[2]

    def py(self):
        return self.rparticles[3]

    def tau(self):
        return self.rparticles[4]

    def p(self):
        return self.rparticles[5]

    def delete_particles(self, inds, record=True):
        """particle.py:323-333 (without the lost-particle recorder)."""
        self.rparticles = np.delete(self.rparticles, inds, 1)
        self.q_array = np.delete(self.q_array, inds, 0)


class DeviceParticleArray:
    """Device-resident particles: ``rparticles`` is a (6, n) view into a
    (6, ld) CUDA fp64 buffer, ``q_array`` a (n,) CUDA fp64 tensor."""

    ALIGN = 16  # doubles (128 bytes)

    def __init__(self, n=0, device=None):
        import torch
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        ld = max(self.ALIGN, (int(n) + self.ALIGN - 1) // self.ALIGN * self.ALIGN)
        self._buf = torch.zeros((6, ld), dtype=torch.float64, device=self.device)
        self._q = torch.zeros(ld, dtype=torch.float64, device=self.device)
        self._n = int(n)
        self.s = 0.0
        self.E = 0.0
        # lost-particle bookkeeping of the reference container (particle.py:24-45)
        self.lost_particles = []
        self.lp_to_pos_hist = []
        self._current_particle = torch.arange(int(n), device=self.device)

    @property
    def n(self):
        return self._n

    def size(self):
        return self._n

    def __len__(self):
        return self._n

    @property
    def rparticles(self):
        return self._buf[:, :self._n]

    @property
    def q_array(self):
        return self._q[:self._n]

    # accessors of the reference container (particle.py:176-195), as device views
    def x(self):
        return self.rparticles[0]

    def px(self):
        return self.rparticles[1]

    def y(self):
        return self.rparticles[2]

    def py(self):
        return self.rparticles[3]

    def tau(self):
        return self.rparticles[4]

    def p(self):
        return self.rparticles[5]

    def delete_particles(self, inds, record=True):
        """Remove particles by index (or boolean mask), keeping the order of the survivors
        (``ParticleArray.delete_particles``, beam/particle.py:323-333).  The buffer is compacted in place."""
        import torch
        inds = torch.as_tensor(inds, device=self.device)
        keep = torch.ones(self._n, dtype=torch.bool, device=self.device)
        if inds.dtype == torch.bool:
            keep &= ~inds
        elif inds.numel():
            keep[inds.long()] = False
        lost = torch.nonzero(~keep).flatten()
        if record:
            self.lost_particles += self._current_particle[lost].tolist()
            self.lp_to_pos_hist.append((self.s, int(lost.numel())))
            self._current_particle = self._current_particle[keep]
        m = int(keep.sum().item())
        if m == self._n:
            return
        self._buf[:, :m] = self._buf[:, :self._n][:, keep]
        self._q[:m] = self._q[:self._n][keep]
        self._n = m

    def cut(self, kind, row, params, record=True):
        """Aperture cut with ordered compaction on the device (``ocl_sc_aperture_cut``: three small kernels, survivors
        keep their order like ``ParticleArray.delete_particles``, beam/particle.py:323-333).  kind 0: lose particles whose
        row ``row`` is outside [params[0], params[1]]; kind 1: outside the ellipse (ax, ay, dx, dy) in (x, y).  One
        host synchronisation, for the new particle count.  Returns the number of lost particles."""
        import torch
        from .beam import _scratch_solver
        if self._n == 0:
            return 0
        ld = self._buf.shape[1]
        new_buf = torch.zeros_like(self._buf)
        new_q = torch.zeros_like(self._q)
        ids = self._current_particle
        if ids.dtype != torch.int64 or not ids.is_contiguous():
            ids = ids.to(torch.int64).contiguous()
        new_ids = torch.empty(self._n, dtype=torch.int64, device=self.device)
        lost = torch.empty(self._n, dtype=torch.int64, device=self.device)
        m = _scratch_solver(self.device.index or 0).aperture_cut(self.rparticles, self.q_array, ids, kind, row, params,
                                                                 new_buf, new_q, new_ids, lost)
        n_lost = self._n - m
        if n_lost == 0:
            return 0
        if record:
            self.lost_particles += lost[:n_lost].tolist()
            self.lp_to_pos_hist.append((self.s, n_lost))
            self._current_particle = new_ids[:m]
        self._buf, self._q, self._n = new_buf, new_q, m
        return n_lost

    @classmethod
    def from_host(cls, p_array, device=None):
        """Copy any object with rparticles/q_array/E/s (e.g. Ocelot's ParticleArray) to the device."""
        import torch
        r = np.ascontiguousarray(p_array.rparticles, dtype=np.float64)
        out = cls(r.shape[1], device=device)
        out.rparticles.copy_(torch.from_numpy(r))
        out.q_array.copy_(torch.from_numpy(np.ascontiguousarray(p_array.q_array, dtype=np.float64)))
        out.E = float(p_array.E)
        out.s = float(getattr(p_array, "s", 0.0))
        return out

    def to_host(self, p_array=None):
        """Copy back into ``p_array`` (in place) or into a new ParticleArray."""
        if p_array is None:
            p_array = ParticleArray(self._n)
        r, q = self.rparticles.cpu().numpy(), self.q_array.cpu().numpy()
        if p_array.rparticles.shape == r.shape:
            p_array.rparticles[:] = r
            p_array.q_array[:] = q
        else:                                   # apertures changed N on the device: new arrays, as the reference's
            p_array.rparticles = r              # delete_particles does on the host (beam/particle.py:323-333)
            p_array.q_array = q
        p_array.E = self.E
        p_array.s = self.s
        return p_array
