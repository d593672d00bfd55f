"""Host-side mirror of the reference plugin protocol.

``PhysProc`` restates the duck-typed base class the reference's ``Navigator`` /
``track()`` drive (ocelot/cpbd/physics_proc.py:14-70): attribute ``step``,
methods ``prepare(lat)``, ``apply(p_array, dz)``, ``finalize()``, and the
attributes the caller injects (navi.py:68-81, track.py:476).  It does not
import Ocelot, so the package also works where Ocelot is not installed.
"""
from __future__ import annotations


class PhysProc:
    """Parent class of physics processes (physics_proc.py:14-70)."""

    def __init__(self, step=1):
        self.step = step          # in units of Navigator.unit_step
        self.energy = None
        self.indx0 = None         # injected by Navigator.add_physics_proc (navi.py:68-78)
        self.indx1 = None
        self.s_start = None
        self.s_stop = None
        self.start_elem = None
        self.end_elem = None
        self.z0 = None            # injected by track() before apply (track.py:476)

    def check_step(self):
        # same predicate as physics_proc.py:41-43
        if not isinstance(self.step, (int, float)) and float(self.step).is_integer():
            raise ValueError(f'step must be an integer number, instead {self.step}')

    def prepare(self, lat):
        self.check_step()

    def apply(self, p_array, dz):
        pass

    def finalize(self, *args, **kwargs):
        pass
