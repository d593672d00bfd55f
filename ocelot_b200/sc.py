"""B200-native drop-in for Ocelot's 3D space-charge physics process.

``SpaceCharge`` keeps the constructor, attributes, ``prepare/apply/finalize``
protocol, ``__repr__`` and side effects of the reference class
(ocelot/cpbd/sc.py:76-258) so ``Navigator.add_physics_proc`` / ``track()``
(ocelot/cpbd/navi.py:63-98, ocelot/cpbd/track.py:470-477) drive it unchanged.
All arithmetic happens in hand-written sm_100a kernels behind the C ABI of
``include/ocelot_sc.h``; this file only marshals arguments.  There is no CPU
path: without the native library or a CUDA device ``apply`` raises.
"""
from __future__ import annotations

import logging

import numpy as np

from .physproc import PhysProc

logger = logging.getLogger(__name__)


class SpaceCharge(PhysProc):
    """Space Charge physics process (API of ocelot/cpbd/sc.py:76-102).

    Attributes:
        step       -- kick every ``step`` Navigator.unit_step
        nmesh_xyz  -- [63, 63, 63] 3D mesh, read at every ``apply`` (users mutate it
                      after construction, space_charge_test.py:71-77)
        random_mesh, random_seed, low_order_kick, debug -- as in the reference
        device     -- CUDA device index used for host-array kicks (default: current)
        deterministic -- extension (default False): deposit the charge in np.bincount's order (sc.py:193) instead
                      of with atomics, so the grid is bit-identical from run to run (a debugging aid, slower)
    """

    def __init__(self, step=1, **kwargs):
        PhysProc.__init__(self)
        self.step = step
        self.nmesh_xyz = kwargs.get("nmesh_xyz", [63, 63, 63])
        self.low_order_kick = kwargs.get("low_order_kick", True)   # stored, never read (sc.py:96)
        self.start_elem = None
        self.end_elem = None
        self.debug = False
        self.random_mesh = kwargs.get("random_mesh", False)
        self.random_seed = 10
        self.device = kwargs.get("device", None)
        self.deterministic = kwargs.get("deterministic", False)    # ordered deposit (debugging aid), see class doc
        self._solvers = {}
        # unknown kwargs are ignored, as in the reference (sc.py:92-102)

    # -- protocol -----------------------------------------------------------
    def prepare(self, lat):
        self.check_step()
        if self.random_seed is not None:
            np.random.seed(self.random_seed)      # global RNG, sc.py:104-107

    def apply(self, p_array, zstep):
        logger.debug(" apply: zstep = %s", zstep)
        if zstep == 0:                             # sc.py:210-212
            return
        r = p_array.rparticles
        n = r.shape[1]
        if n == 0:
            return
        draws = None
        if self.random_mesh:                       # draw order of sc.py:175, :185
            draws = (np.random.uniform(low=1, high=1.1), np.random.uniform(low=-0.5, high=0.5))
        E = float(p_array.E)
        if isinstance(r, np.ndarray):
            solver = self._solver(self._host_device())
            solver.kick_host(r, p_array.q_array, E, float(zstep), draws)
        else:                                      # device-resident torch tensors
            dev = r.device.index if r.device.index is not None else 0
            self._solver(dev).kick_device(r, p_array.q_array, E, float(zstep), draws)

    # -- native handle management ------------------------------------------
    def _host_device(self):
        if self.device is not None:
            return int(self.device)
        import torch
        if not torch.cuda.is_available():
            raise RuntimeError("ocelot_b200.SpaceCharge needs a CUDA device; there is no CPU fallback")
        return torch.cuda.current_device()

    def _solver(self, device):
        from . import native
        nmesh = tuple(int(v) for v in np.array(self.nmesh_xyz).reshape(-1))
        key = (int(device), nmesh)
        s = self._solvers.get(key)
        if s is None:
            s = native.Solver(int(device), nmesh)
            self._solvers = {key: s}               # one live handle; a mesh change re-plans
        if getattr(self, "deterministic", False) != getattr(s, "_ordered", False):
            s.set_deterministic(bool(self.deterministic))
            s._ordered = bool(self.deterministic)
        return s

    # Navigator deep-copies the process table (navi.py:189) and ParameterScanner
    # pickles navigators (track.py:682-708): drop native handles, recreate lazily.
    def __getstate__(self):
        state = self.__dict__.copy()
        state["_solvers"] = {}
        return state

    def __setstate__(self, state):
        self.__dict__.update(state)
        self._solvers = {}

    def __deepcopy__(self, memo):
        import copy
        new = self.__class__.__new__(self.__class__)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            new.__dict__[k] = {} if k == "_solvers" else copy.deepcopy(v, memo)
        return new

    def __repr__(self) -> str:
        cname = type(self).__name__
        step = self.step
        nmesh_xyz = self.nmesh_xyz
        random_mesh = self.random_mesh
        return f"<{cname}: {step=}, {nmesh_xyz=}, {random_mesh=}>"


def install():
    """Replace the reference ``SpaceCharge`` with this class in every module that already holds it
    (``ocelot.cpbd.sc``, the ``ocelot`` re-export, and any module that star-imported it, e.g.
    ``ocelot.utils.section_track``), so unmodified Ocelot scripts pick up the B200 kick."""
    from ._install import swap
    return swap("SpaceCharge", SpaceCharge)


def uninstall():
    """Undo ``install()``."""
    from ._install import restore
    restore("SpaceCharge")
