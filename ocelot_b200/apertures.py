"""Apertures for host or device-resident bunches: the only in-loop processes that change the
number of particles between kicks.  Same constructors, attributes and selection rules as
``RectAperture`` / ``EllipticalAperture`` (ocelot/cpbd/physics_proc.py:341-390); on a
DeviceParticleArray the mask is evaluated and the arrays compacted on the GPU."""
from __future__ import annotations

import numpy as np

from .physproc import PhysProc


def _lib(x):
    if isinstance(x, np.ndarray):
        return np
    import torch
    return torch


class RectAperture(PhysProc):
    """Cut the beam in the horizontal and/or vertical plane (physics_proc.py:341-371)."""
    device_resident = True

    def __init__(self, xmin=-np.inf, xmax=np.inf, ymin=-np.inf, ymax=np.inf, step=1):
        PhysProc.__init__(self, step)
        self.xmin, self.xmax, self.ymin, self.ymax = xmin, xmax, ymin, ymax

    def apply(self, p_array, dz):
        if hasattr(p_array, "cut"):                       # device-resident: predicate + ordered compaction kernels
            p_array.cut(0, 0, (self.xmin, self.xmax, 0.0, 0.0))                                # :362-365
            p_array.cut(0, 2, (self.ymin, self.ymax, 0.0, 0.0))                                # :367-370
            return
        x = p_array.x()
        lib = _lib(x)
        p_array.delete_particles(_where(lib, lib.logical_or(x < self.xmin, x > self.xmax)))   # :362-365
        y = p_array.y()
        p_array.delete_particles(_where(lib, lib.logical_or(y < self.ymin, y > self.ymax)))   # :367-370


class EllipticalAperture(PhysProc):
    """Delete particles outside an ellipse (physics_proc.py:373-390)."""
    device_resident = True

    def __init__(self, xmax=np.inf, ymax=None, dx=0.0, dy=0.0, step=1):
        PhysProc.__init__(self, step)
        self.xmax = xmax
        self.ymax = ymax if ymax is not None else xmax
        self.dx, self.dy = dx, dy

    def apply(self, p_array, dz):
        if hasattr(p_array, "cut"):                       # device-resident: predicate + ordered compaction kernels
            p_array.cut(1, 0, (self.xmax, self.ymax, self.dx, self.dy))
            return
        x, y = p_array.x(), p_array.y()
        lib = _lib(x)
        out = (x - self.dx) ** 2 / self.xmax ** 2 + (y - self.dy) ** 2 / self.ymax ** 2 > 1.0   # :388
        p_array.delete_particles(_where(lib, out))


def _where(lib, mask):
    if lib is np:
        return np.argwhere(mask).reshape(-1)
    return lib.nonzero(mask).flatten()
