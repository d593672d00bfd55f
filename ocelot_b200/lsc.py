"""B200-native drop-in for Ocelot's longitudinal space-charge process ``LSC``.

Same constructor, attributes and ``prepare/apply/finalize`` protocol as the reference class
(ocelot/cpbd/sc.py:261-599), so ``Navigator.add_physics_proc`` / ``track()`` drive it unchanged.
Every per-particle operation (bunch statistics, the first-order current deposit, the slice
size, the wake interpolation and the energy kick) and the 1-D impedance solve run in the sm_100a
kernels of ``csrc/sc_lsc.cu`` behind ``ocl_sc_lsc_*`` (include/ocelot_sc.h).  The host only
derives a dozen scalars per kick -- the grid definition of ``s_to_cur``
(ocelot/cpbd/beam/analysis.py:293-333) and the undulator factor -- with the reference's own
float expressions.  There is no CPU path: without the native library or a CUDA device
``apply`` raises.
"""
from __future__ import annotations

import logging

import numpy as np

from .constants import m_e_GeV, speed_of_light
from .physproc import PhysProc

logger = logging.getLogger(__name__)


def _is_undulator(elem) -> bool:
    """The reference tests ``isinstance(elem, Undulator)`` (sc.py:502); this package does not
    import Ocelot, so the class is recognised by name anywhere in its MRO."""
    return any(c.__name__ == "Undulator" for c in type(elem).__mro__)


class LSC(PhysProc):
    """Longitudinal Space Charge impedance model (API of ocelot/cpbd/sc.py:261-297).

    Attributes:
        step          -- kick every ``step`` Navigator.unit_step
        step_profile  -- uniform transverse profile (imp_step_lsc) instead of round Gaussian (imp_lsc)
        smooth_param  -- current-profile smoothing: resolution = std(tau) * smooth_param
        bounds        -- [min, max] in units of std(tau): central slice that defines the beam size
        slice         -- stored, unused (as in the reference)
        device        -- CUDA device index used for host-array kicks (default: current)
        async_grid    -- True (default): the 1-D grid is derived on the device from the bunch statistics, so a
                         kick involves no host synchronisation (grids of up to 8192 points); False: the host
                         derives it after one small device->host read (any grid size)
    """

    def __init__(self, step=1, **kwargs):
        PhysProc.__init__(self, step)
        self.K_s_func = None
        self.step_profile = kwargs.get("step_profile", False)
        self.smooth_param = kwargs.get("smooth_param", 0.1)
        self.bounds = kwargs.get("bounds", [-0.4, 0.4])
        self.slice = kwargs.get("slice", None)
        self._is_undul_in_beam_line = False
        self.device = kwargs.get("device", None)
        self.async_grid = kwargs.get("async_grid", True)
        self._solvers = {}
        self._stage = {}
        self._async_verified = False
        self._last_params = None
        self._last_solver = None

    @property
    def last_params(self):
        """Scalars of the last kick (grid a, ds, nb, taps K, slice bounds, charge ...).  After an asynchronous
        kick they are read back from the device on first use."""
        if self._last_params is None and self._last_solver is not None:
            self._last_params = self._last_solver.lsc_last_params()
        return self._last_params

    @last_params.setter
    def last_params(self, value):
        self._last_params = value

    # -- protocol -----------------------------------------------------------
    def prepare(self, lat):
        """Piecewise-constant undulator strength K(s) over [start_elem, end_elem], as a linear interpolant
        through both edges of every element (what sc.py:476-508 builds): K = Kx + Ky inside planar
        undulators (at most one of Kx, Ky non-zero), 0 elsewhere, end values held outside the range."""
        from scipy.interpolate import interp1d
        self.check_step()
        elems = list(lat.get_sequence_part(self.start_elem, self.end_elem))
        # element edges by sequential addition from s_start (same rounding as a running sum)
        edges = np.cumsum([self.s_start] + [e.l for e in elems])
        strength = []
        for e in elems:
            planar = _is_undulator(e) and not (e.Kx != 0 and e.Ky != 0)
            strength.append(e.Kx + e.Ky if planar else 0.0)
            self._is_undul_in_beam_line = self._is_undul_in_beam_line or planar
        s = np.repeat(edges, 2)[1:-1]               # entry and exit edge of every element
        k = np.repeat(np.asarray(strength, dtype=float), 2)
        self.K_s_func = interp1d(s, k, kind='linear', bounds_error=False, fill_value=(k[0], k[-1]))

    def compute_filling_factor(self, x0, x1, num_points=100):
        """Fraction of [x0, x1] with a non-zero K at either end of a sampling interval (sc.py:510-544)."""
        on = self.K_s_func(np.linspace(x0, x1, num_points)) != 0
        covered = np.sum(on[:-1] | on[1:]) * ((x1 - x0) / (num_points - 1))
        return covered / (x1 - x0)

    def undulator_factor(self, dz):
        """(K_max, fill_factor) of the step that ends at z0 (sc.py:569-574)."""
        if self._is_undul_in_beam_line:
            fill_factor = self.compute_filling_factor(self.z0 - dz, self.z0, num_points=200)
            K_max = np.max(self.K_s_func(np.linspace(self.z0 - dz, self.z0, num=200)))
        else:
            fill_factor = 0
            K_max = 0
        return K_max, fill_factor

    def kick_parameters(self, stats: dict, E, dz) -> dict:
        """The scalars of one kick, from the bunch statistics of sweep A, with the reference's own
        expressions: slice bounds (sc.py:579-580), charge and velocity (:590-592), the grid of
        s_to_cur (analysis.py:293-321) and its smoothing taps (:330-331)."""
        n = stats["n"]
        mean_tau = stats["mean_tau"]
        sigma_tau = np.sqrt(stats["m2_tau"] / n)                       # np.std
        K_max, fill_factor = self.undulator_factor(dz)
        q = stats["sum_q"]
        gamma = E / m_e_GeV
        v = np.sqrt(1 - 1 / gamma ** 2) * speed_of_light
        sigma = sigma_tau * self.smooth_param
        Nsigma = 3
        a = stats["min_tau"] - Nsigma * sigma
        b = stats["max_tau"] + Nsigma * sigma
        if sigma > 0:
            ds = 0.25 * sigma
        else:
            ds = (b - a) / 1000.0
        N = int(np.ceil((b - a) / ds))
        ds = (b - a) / N
        K = int(np.floor(Nsigma * sigma / ds + 0.5)) if sigma > 0 else -1
        return dict(
            slice_min=mean_tau + sigma_tau * self.bounds[0], slice_max=mean_tau + sigma_tau * self.bounds[1],
            x_shift=stats["sum_x"] / n, y_shift=stats["sum_y"] / n,
            a=a, ds=ds, nb=N + 1, sigma_s=sigma, K=K, q=q, v=v, gamma=gamma, dz=dz,
            und=1 + 0.5 * K_max * K_max * fill_factor,
            pc_ref=np.sqrt(E ** 2 / m_e_GeV ** 2 - 1) * m_e_GeV,
            step_profile=1.0 if self.step_profile else 0.0, n_total=n)

    def apply(self, p_array, dz):
        if dz < 1e-10:                                                   # sc.py:566-568
            logger.debug(" LSC applied, dz < 1e-10, dz = " + str(dz))
            return
        r = p_array.rparticles
        if r.shape[1] == 0:
            return
        E = float(p_array.E)
        if isinstance(r, np.ndarray):
            self._apply_host(r, p_array.q_array, E, float(dz))
        else:
            dev = r.device.index if r.device.index is not None else 0
            self._kick_device(self._solver(dev), r, p_array.q_array, E, float(dz))

    def _apply_host(self, r, q_array, E, dz):
        """Host numpy arrays: stage the four rows LSC reads (x, y, tau, delta) to the device through the
        handle-owned pinned staging buffers, kick, and copy row 5 back in place (the only row LSC writes,
        sc.py:599)."""
        import torch
        dev = self._host_device()
        solver = self._solver(dev)
        n = r.shape[1]
        with torch.cuda.device(dev):
            st = self._staging(dev, n)
            d_r, d_q, pin = st["d_r"][:, :n], st["d_q"][:n], st["pin"]
            for slot, row in enumerate((0, 2, 4, 5)):
                pin[slot, :n].copy_(torch.from_numpy(r[row]))                 # host -> pinned (memcpy rate)
                d_r[row].copy_(pin[slot, :n], non_blocking=True)              # pinned -> device (PCIe rate)
            pin[4, :n].copy_(torch.from_numpy(np.ascontiguousarray(q_array, dtype=np.float64)))
            d_q.copy_(pin[4, :n], non_blocking=True)
            self._kick_device(solver, d_r, d_q, E, dz, checked=True)
            pin[3, :n].copy_(d_r[5], non_blocking=True)
            torch.cuda.current_stream().synchronize()
            r[5][:] = pin[3, :n].numpy()

    def _staging(self, dev, n):
        import torch
        st = self._stage.get(dev)
        if st is None or st["cap"] < n:
            cap = max(n + n // 8, 1024)
            st = dict(cap=cap, d_r=torch.empty((6, cap), dtype=torch.float64, device=f"cuda:{dev}"),
                      d_q=torch.empty(cap, dtype=torch.float64, device=f"cuda:{dev}"),
                      pin=torch.empty((5, cap), dtype=torch.float64).pin_memory())
            self._stage = {dev: st}
        return st

    def _kick_device(self, solver, r, q, E, dz, checked=False):
        """``checked``: the caller synchronises anyway (host arrays), so the outcome of the asynchronous form is
        read right away.  Otherwise the FIRST asynchronous kick of this object is checked once (one stream
        synchronisation per tracking run); a grid that does not fit the asynchronous form (more than 8192 points:
        small ``smooth_param``, long tails) switches the object to the host-derived grid for good and the kick
        is redone -- never skipped.  Later overflows surface in the next ``apply`` / ``finalize`` as an error."""
        if self.async_grid:
            late = solver.lsc_async_status(synchronise=False)
            if late:
                self.async_grid = False
                raise RuntimeError("ocelot_b200.LSC: an earlier asynchronous kick was skipped on the device (its grid "
                                   "outgrew the asynchronous form); set async_grid=False for this bunch")
            K_max, fill_factor = self.undulator_factor(dz)
            gamma = E / m_e_GeV
            solver.lsc_kick_async(r, q, gamma, np.sqrt(1 - 1 / gamma ** 2) * speed_of_light,
                                  np.sqrt(E ** 2 / m_e_GeV ** 2 - 1) * m_e_GeV, dz,
                                  1 + 0.5 * K_max * K_max * fill_factor, self.bounds, self.smooth_param,
                                  self.step_profile)
            self._last_params, self._last_solver = None, solver
            if checked or not self._async_verified:
                self._async_verified = True
                if solver.lsc_async_status(synchronise=True):
                    logger.warning("LSC: the current-profile grid does not fit the asynchronous form; "
                                   "falling back to the host-derived grid (async_grid=False)")
                    self.async_grid = False                     # the skipped kick left the particles untouched
                else:
                    return
            else:
                return
        params = self.kick_parameters(solver.lsc_stats(r, q), E, dz)
        solver.lsc_kick(r, params)
        self._last_params, self._last_solver = params, None

    def finalize(self, *args, **kwargs):
        """End of a tracking run (track.py:498-499): the last asynchronous kick must not be dropped silently."""
        for solver in self._solvers.values():
            if solver.lsc_async_status(synchronise=True):
                raise RuntimeError("ocelot_b200.LSC: the last asynchronous kick was skipped on the device (its grid "
                                   "outgrew the asynchronous form); rerun with async_grid=False")

    # -- host-side utilities of the reference class (plotting / analysis helpers; ``apply`` does not
    #    use them: the kick evaluates the same formulas on the device, csrc/sc_lsc.cu) -------------
    def imp_lsc(self, gamma, sigma, w, dz):
        """Round-Gaussian-beam LSC impedance Z(w) [Ohm] over a length dz (sc.py:299-340)."""
        from scipy.special import exp1, factorial
        from .constants import epsilon_0, pi
        Z0 = 1. / (speed_of_light * epsilon_0)
        alpha2 = (w * sigma / (gamma * speed_of_light)) ** 2
        T = np.zeros(w.shape)
        mid = (alpha2 <= 40.0) & (alpha2 >= 1e-16)
        far = alpha2 > 40.0
        T[mid] = np.exp(alpha2[mid]) * exp1(alpha2[mid])
        T[far] = sum((-1) ** i * factorial(i) / alpha2[far] ** (i + 1) for i in range(10))
        return 1j * Z0 / (4 * pi * speed_of_light * gamma ** 2) * w * T * dz

    def imp_step_lsc(self, gamma, rb, w, dz):
        """Uniform-beam LSC impedance (sc.py:342-369); clamps ``w`` in place like the reference."""
        from scipy.special import k1
        from .constants import epsilon_0
        Z0 = 1. / (speed_of_light * epsilon_0)
        low = np.where(w < 1e-7)[0]
        w[low] = 1e-7
        x = w * rb / (speed_of_light * gamma)
        Z = 1j * Z0 * speed_of_light / (4 * w * rb * rb) * dz * (1 - x * k1(x))
        Z[low] = 0
        return Z

    def wake2impedance(self, s, w):
        """Spectrum of a signal on a uniform grid, exp(+iwt) convention (sc.py:371-398,
        beam/analysis.py:343-383): returns (f [Hz], y)."""
        dt = (s[1] - s[0]) / speed_of_light
        n = len(s)
        return 1 / dt * np.arange(0, n) / n, dt * np.fft.fft(w, n)

    def impedance2wake(self, f, y):
        """Inverse of wake2impedance for a Hermitian spectrum (sc.py:401-416): returns (s [m], w)."""
        df = f[1] - f[0]
        n = len(f)
        return 1 / df * np.arange(0, n) / n * speed_of_light, n * df * np.fft.irfft(y, n)

    # -- native handle management (as SpaceCharge) ------------------------------
    def _host_device(self):
        if self.device is not None:
            return int(self.device)
        import torch
        if not torch.cuda.is_available():
            raise RuntimeError("ocelot_b200.LSC needs a CUDA device; there is no CPU fallback")
        return torch.cuda.current_device()

    def _solver(self, device):
        from . import native
        s = self._solvers.get(int(device))
        if s is None:
            s = native.Solver(int(device), (4, 4, 4))      # LSC uses none of the 3-D grids
            self._solvers = {int(device): s}
        return s

    def __getstate__(self):
        state = self.__dict__.copy()
        state["_solvers"] = {}
        state["_stage"] = {}
        state["_last_solver"] = None
        return state

    def __setstate__(self, state):
        self.__dict__.update(state)
        self._solvers = {}
        self._stage = {}

    def __deepcopy__(self, memo):
        import copy
        new = self.__class__.__new__(self.__class__)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            new.__dict__[k] = {} if k in ("_solvers", "_stage") else (None if k == "_last_solver" else copy.deepcopy(v, memo))
        return new


def install():
    """Replace the reference ``LSC`` with this class in every module that already holds it."""
    from ._install import swap
    return swap("LSC", LSC)


def uninstall():
    """Undo ``install()``."""
    from ._install import restore
    restore("LSC")
