"""ctypes binding of the C-ABI library (include/ocelot_sc.h).

The library is the product: there is no Python or CPU fallback.  If
``libocelot_sc.so`` is absent it is built once with nvcc; if that is impossible,
or if no CUDA device is present when a solver is created, a ``RuntimeError`` is
raised.
"""
from __future__ import annotations

import ctypes as C
import os
import re
import weakref

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("OCL_SC_LIB") or os.path.join(_HERE, "libocelot_sc.so")   # OCL_SC_LIB: A/B builds
HEADER = os.path.join(os.path.dirname(_HERE), "include", "ocelot_sc.h")

BUF_MOMENTUM, BUF_EXTENT_MAX, BUF_EXTENT_SUM, BUF_RHO, BUF_EXTENT = 0, 1, 2, 3, 4
BUF_RHO_SLAB, BUF_XCHG_A, BUF_XCHG_B, BUF_PHI_SLAB, BUF_PHI = 5, 6, 7, 8, 9
BUF_LSC_BINS, BUF_LSC_SLICE_MAX, BUF_LSC_SLICE_SUM = 10, 11, 12

_lib = None

_dp = C.POINTER(C.c_double)
_vp = C.c_void_p
_ll = C.c_longlong

_SIGNATURES = {
    "ocl_sc_abi_version": (C.c_int, []),
    "ocl_sc_get_constants": (None, [_dp]),
    "ocl_sc_fft_size": (C.c_int, [C.c_int]),
    "ocl_sc_create": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, _ll, C.POINTER(_vp)]),
    "ocl_sc_destroy": (None, [_vp]),
    "ocl_sc_last_error": (C.c_char_p, [_vp]),
    "ocl_sc_kick_device": (C.c_int, [_vp, _vp, _ll, _vp, _ll, C.c_double, C.c_double, _dp, _vp]),
    "ocl_sc_kick_host": (C.c_int, [_vp, _vp, _ll, _vp, _ll, C.c_double, C.c_double, _dp]),
    "ocl_sc_host_register": (C.c_int, [_vp, _ll]),
    "ocl_sc_host_unregister": (C.c_int, [_vp]),
    "ocl_sc_collective_buffer": (C.c_int, [_vp, C.c_int, C.POINTER(_vp), C.POINTER(_ll)]),
    "ocl_sc_combine_extents": (C.c_int, [_vp, _vp, C.c_int, _vp]),
    "ocl_sc_slab_init": (C.c_int, [_vp, C.c_int, C.c_int]),
    "ocl_sc_slab_forward": (C.c_int, [_vp, _vp]),
    "ocl_sc_slab_xpass": (C.c_int, [_vp, _vp]),
    "ocl_sc_slab_inverse": (C.c_int, [_vp, _vp]),
    "ocl_sc_slab_finish": (C.c_int, [_vp, _dp, _vp]),
    "ocl_sc_mailbox_init": (C.c_int, [_vp, C.c_int, C.c_int, C.POINTER(_vp)]),
    "ocl_sc_mailbox_exchange": (C.c_int, [_vp, C.c_int, _vp]),
    "ocl_sc_mailbox_status": (C.c_int, [_vp, C.c_int, C.POINTER(C.c_int)]),
    "ocl_sc_set_peer_rho": (C.c_int, [_vp, C.c_int, C.c_int, C.POINTER(_vp)]),
    "ocl_sc_set_peer_xchg": (C.c_int, [_vp, C.c_int, C.c_int, C.POINTER(_vp), C.POINTER(_vp)]),
    "ocl_sc_set_multicast_rho": (C.c_int, [_vp, _vp, _vp]),
    "ocl_sc_set_multicast_phi": (C.c_int, [_vp, _vp, _vp]),
    "ocl_sc_nvls_reduce_rho": (C.c_int, [_vp, _vp]),
    "ocl_sc_use_device_params": (C.c_int, [_vp, C.c_int]),
    "ocl_sc_set_kick_params": (C.c_int, [_vp, C.c_double, C.c_double, _dp, _vp]),
    "ocl_sc_stage_momentum": (C.c_int, [_vp, _vp, _ll, _ll, C.c_double, _vp]),
    "ocl_sc_stage_extent": (C.c_int, [_vp, _vp, _ll, _vp, _ll, C.c_double, _dp, _vp]),
    "ocl_sc_defer_finish": (C.c_int, [_vp, C.c_int]),
    "ocl_sc_set_deterministic": (C.c_int, [_vp, C.c_int]),
    "ocl_sc_stage_finish": (C.c_int, [_vp, C.c_int, C.c_double, _dp, _vp]),
    "ocl_sc_stage_deposit": (C.c_int, [_vp, _vp, _ll, _vp, _ll, C.c_double, _dp, _vp]),
    "ocl_sc_stage_solve": (C.c_int, [_vp, _dp, _vp]),
    "ocl_sc_stage_kick": (C.c_int, [_vp, _vp, _ll, _ll, C.c_double, C.c_double, _dp, _vp]),
    "ocl_sc_get_geometry": (C.c_int, [_vp, _dp]),
    "ocl_sc_get_rho": (C.c_int, [_vp, _vp]),
    "ocl_sc_get_phi": (C.c_int, [_vp, _vp]),
    "ocl_sc_get_green": (C.c_int, [_vp, _vp]),
    "ocl_sc_field_at_particles": (C.c_int, [_vp, _vp, _ll, _vp, _ll, C.c_double, _dp, _vp, _vp]),
    "ocl_sc_mad_to_cartesian": (C.c_int, [_vp, _vp, _ll, _ll, C.c_double, _vp, _ll, _vp]),
    "ocl_sc_cartesian_to_mad": (C.c_int, [_vp, _vp, _ll, _ll, C.c_double, _vp, _ll, _vp]),
    "ocl_sc_potential_host": (C.c_int, [_vp, _vp, _dp, _vp]),
    "ocl_sc_map_apply": (C.c_int, [_vp, _vp, _ll, _ll, _dp, _dp, _dp, _vp]),
    "ocl_sc_cavity_apply": (C.c_int, [_vp, _vp, _ll, _ll, _dp, _dp, _dp, C.c_int, _vp]),
    "ocl_sc_beam_moments": (C.c_int, [_vp, _vp, _ll, _ll, _dp, _vp]),
    "ocl_sc_beam_moments_device": (C.c_int, [_vp, _vp, _ll, _ll, _vp, _vp, _vp]),
    "ocl_sc_aperture_cut": (C.c_int, [_vp, _vp, _ll, _vp, _vp, _ll, C.c_int, C.c_int, _dp, _vp, _ll, _vp, _vp, _vp,
                                      C.POINTER(_ll), _vp]),
    "ocl_sc_cavity_coefficients": (C.c_int, [C.c_double] * 6 + [_dp, C.POINTER(C.c_int), _dp]),
    "ocl_sc_lsc_stats": (C.c_int, [_vp, _vp, _ll, _ll, _vp, _dp, _vp]),
    "ocl_sc_lsc_kick": (C.c_int, [_vp, _vp, _ll, _ll, _dp, _vp]),
    "ocl_sc_lsc_deposit": (C.c_int, [_vp, _vp, _ll, _ll, _dp, _vp]),
    "ocl_sc_lsc_solve_kick": (C.c_int, [_vp, _vp, _ll, _ll, _dp, _vp]),
    "ocl_sc_lsc_get_profile": (C.c_int, [_vp, C.c_int, _dp, _dp, _dp]),
    "ocl_sc_lsc_kick_async": (C.c_int, [_vp, _vp, _ll, _ll, _vp, _dp, _vp]),
    "ocl_sc_lsc_last_params": (C.c_int, [_vp, _dp]),
    "ocl_sc_lsc_async_status": (C.c_int, [_vp, C.c_int, C.POINTER(C.c_int)]),
    "ocl_sc_enable_timers": (C.c_int, [_vp, C.c_int]),
    "ocl_sc_get_timers": (C.c_int, [_vp, _dp]),
    "ocl_sc_launch_count": (_ll, [_vp]),
}


def declared_symbols() -> list[str]:
    """Names of the functions include/ocelot_sc.h declares."""
    with open(HEADER) as f:
        text = f.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ocl_sc_[a-z_0-9]+)\s*\(", text)))


def load():
    """Load (building if needed) the native library; raises if impossible."""
    global _lib
    if _lib is not None:
        return _lib
    from . import build as _build
    exists = os.path.exists(LIB_PATH)
    stale = (exists and LIB_PATH == _build.LIB and not _build.up_to_date()
             and os.access(os.path.dirname(LIB_PATH), os.W_OK) and _build.have_nvcc())
    if not exists or stale:                        # stale: the sources differ from what the library was built from
        try:
            _build.build(force=stale)
        except Exception as exc:  # noqa: BLE001
            if exists:                             # keep running on the library that is there; say so
                import warnings
                warnings.warn(f"ocelot_b200: {LIB_PATH} is older than its sources and could not be rebuilt ({exc}); "
                              "using the existing library")
            else:
                raise RuntimeError(
                    f"ocelot_b200: native library {LIB_PATH} is missing and could not be built ({exc}). "
                    "There is no CPU fallback for the space-charge kick.") from exc
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


# ---- page-locking of caller-owned numpy buffers, tied to the owning array's lifetime ----
_pinned = {}          # address of the owning array's buffer -> (bytes, finalizer)
PIN_HOST = os.environ.get("OCL_SC_PIN_HOST", "1") != "0"


def _unpin(ptr):
    _pinned.pop(ptr, None)
    try:
        load().ocl_sc_host_unregister(ptr)
    except Exception:  # noqa: BLE001  (interpreter shutdown)
        pass


def pin_host_array(arr: np.ndarray) -> bool:
    """Page-lock the buffer behind ``arr`` (so host<->device copies run at PCIe rate) for as long as the array
    that OWNS the buffer lives: a ``weakref.finalize`` on the owner unregisters the range before numpy frees it,
    so a later allocation at the same address never inherits a stale registration.  Buffers owned by foreign
    objects (e.g. a pinned torch tensor viewed through ``.numpy()``) are left alone."""
    if not PIN_HOST:
        return False
    owner = arr
    while isinstance(owner.base, np.ndarray):
        owner = owner.base
    if owner.base is not None or not owner.flags.owndata or owner.nbytes == 0:
        return False
    ptr = owner.ctypes.data
    ent = _pinned.get(ptr)
    if ent is not None:
        if ent[0] == owner.nbytes and ent[1].alive:
            return True
        ent[1]()                                   # different extent at this address: drop the old registration
    if load().ocl_sc_host_register(ptr, owner.nbytes) != 0:
        return False
    _pinned[ptr] = (owner.nbytes, weakref.finalize(owner, _unpin, ptr))
    return True


def constants() -> dict:
    out = (C.c_double * 5)()
    load().ocl_sc_get_constants(out)
    return dict(m_e_eV=out[0], m_e_GeV=out[1], epsilon_0=out[2], pi=out[3], speed_of_light=out[4])


def cavity_coefficients(v, phi_deg, freq, E, delta_length, length):
    """(mode, coef[7], delta_e) of the RF-cavity body map (ocl_sc_cavity_coefficients)."""
    coef = (C.c_double * 7)()
    mode, de = C.c_int(0), C.c_double(0.0)
    dl = float("nan") if delta_length is None else float(delta_length)
    rc = load().ocl_sc_cavity_coefficients(float(v), float(phi_deg), float(freq), float(E), dl, float(length),
                                           coef, C.byref(mode), C.byref(de))
    if rc != 0:
        raise RuntimeError("ocl_sc_cavity_coefficients failed")
    return int(mode.value), list(coef), float(de.value)


def fft_size(n: int) -> int:
    return int(load().ocl_sc_fft_size(int(n)))


def _draws(mesh_draws):
    if mesh_draws is None:
        return None
    arr = (C.c_double * 2)(float(mesh_draws[0]), float(mesh_draws[1]))
    return arr


def _stream_ptr(stream):
    if stream is None:
        import torch
        return torch.cuda.current_stream().cuda_stream
    if hasattr(stream, "cuda_stream"):
        return stream.cuda_stream
    return int(stream)


class Solver:
    """One native solver handle = one device + one mesh size."""

    def __init__(self, device: int, nmesh_xyz, max_particles: int = 0):
        self._lib = load()
        self.device = int(device)
        self.nmesh = tuple(int(v) for v in nmesh_xyz)
        if len(self.nmesh) != 3:
            raise ValueError("nmesh_xyz must have three entries")
        h = _vp()
        rc = self._lib.ocl_sc_create(self.device, *self.nmesh, int(max_particles), C.byref(h))
        if rc != 0:
            raise RuntimeError("ocl_sc_create failed: " + self._lib.ocl_sc_last_error(None).decode())
        self._h = h

    # -- lifetime -----------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            self._lib.ocl_sc_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass

    def _check(self, rc, what):
        if rc != 0:
            raise RuntimeError(f"{what} failed: " + self._lib.ocl_sc_last_error(self._h).decode())

    # -- kicks --------------------------------------------------------------
    @staticmethod
    def _dev_rows(r, q=None):
        """(ptr, ld, n) of a torch CUDA fp64 tensor of shape (6, n) with contiguous rows."""
        import torch
        if not (isinstance(r, torch.Tensor) and r.is_cuda and r.dtype == torch.float64 and r.dim() == 2
                and r.shape[0] == 6 and r.stride(1) == 1):
            raise TypeError("rparticles must be a CUDA float64 tensor of shape (6, n) with unit column stride")
        n = r.shape[1]
        if q is not None:
            if not (q.is_cuda and q.dtype == torch.float64 and q.dim() == 1 and q.shape[0] == n and q.stride(0) == 1):
                raise TypeError("q_array must be a contiguous CUDA float64 tensor of length n")
        return r.data_ptr(), (r.stride(0) if n > 0 else 0), n

    def kick_device(self, r, q, E_GeV, dz, mesh_draws=None, stream=None):
        ptr, ld, n = self._dev_rows(r, q)
        self._check(self._lib.ocl_sc_kick_device(self._h, ptr, ld, q.data_ptr(), n, float(E_GeV), float(dz),
                                                 _draws(mesh_draws), _stream_ptr(stream)), "ocl_sc_kick_device")

    def kick_host(self, r: np.ndarray, q: np.ndarray, E_GeV, dz, mesh_draws=None):
        """Kick numpy particles in place.  A writable C-layout float64 (6, n) array is kicked where it lies (and
        page-locked for the lifetime of its owner); anything else numpy's own code would accept (Fortran order,
        slices with a column stride, float32 ...) goes through a contiguous float64 copy that is written back."""
        if not isinstance(r, np.ndarray) or r.ndim != 2 or r.shape[0] != 6:
            raise TypeError("rparticles must be a numpy array of shape (6, n)")
        if not r.flags.writeable:
            raise ValueError("rparticles is read-only: the kick works in place")
        n = r.shape[1]
        direct = r.dtype == np.float64 and (n == 0 or (r.strides[1] == 8 and r.strides[0] % 8 == 0 and r.strides[0] >= 8 * n))
        work = r if direct else np.ascontiguousarray(r, dtype=np.float64)
        qc = np.ascontiguousarray(q, dtype=np.float64)
        if qc.shape != (n,):
            raise ValueError("q_array must have one charge per particle")
        if direct:
            pin_host_array(work)
        if qc is q:
            pin_host_array(qc)                     # a temporary copy is not worth registering
        ld = work.strides[0] // 8 if n > 0 else 0
        self._check(self._lib.ocl_sc_kick_host(self._h, work.ctypes.data, ld, qc.ctypes.data, n, float(E_GeV),
                                               float(dz), _draws(mesh_draws)), "ocl_sc_kick_host")
        if not direct:
            r[...] = work

    # -- stages (sharded operation) ----------------------------------------
    def collective_buffer(self, which: int):
        """torch view (no copy) of one of the handle's small reduction buffers."""
        import torch
        ptr, cnt = _vp(), _ll()
        self._check(self._lib.ocl_sc_collective_buffer(self._h, which, C.byref(ptr), C.byref(cnt)),
                    "ocl_sc_collective_buffer")
        return _wrap_device_doubles(ptr.value, cnt.value, self.device)

    def slab_init(self, rank, world):
        self._check(self._lib.ocl_sc_slab_init(self._h, int(rank), int(world)), "ocl_sc_slab_init")

    def slab_forward(self, stream=None):
        self._check(self._lib.ocl_sc_slab_forward(self._h, _stream_ptr(stream)), "ocl_sc_slab_forward")

    def slab_xpass(self, stream=None):
        self._check(self._lib.ocl_sc_slab_xpass(self._h, _stream_ptr(stream)), "ocl_sc_slab_xpass")

    def slab_inverse(self, stream=None):
        self._check(self._lib.ocl_sc_slab_inverse(self._h, _stream_ptr(stream)), "ocl_sc_slab_inverse")

    def slab_finish(self, mesh_draws=None, stream=None):
        self._check(self._lib.ocl_sc_slab_finish(self._h, _draws(mesh_draws), _stream_ptr(stream)),
                    "ocl_sc_slab_finish")

    def combine_extents(self, gathered, world, stream=None):
        self._check(self._lib.ocl_sc_combine_extents(self._h, gathered.data_ptr(), int(world), _stream_ptr(stream)),
                    "ocl_sc_combine_extents")

    MAILBOX_DOUBLES = 256

    def mailbox_init(self, rank, world, peer_ptrs):
        arr = (_vp * int(world))(*[int(p) for p in peer_ptrs])
        self._check(self._lib.ocl_sc_mailbox_init(self._h, int(rank), int(world), arr), "ocl_sc_mailbox_init")

    def set_peer_rho(self, rank, world, peer_ptrs):
        arr = (_vp * int(world))(*[int(p) for p in peer_ptrs])
        self._check(self._lib.ocl_sc_set_peer_rho(self._h, int(rank), int(world), arr), "ocl_sc_set_peer_rho")

    def set_peer_xchg(self, rank, world, peer_a, peer_b):
        """Slab mode: the y and x passes store straight into the peers' exchange buffers (no all-to-all)."""
        a = (_vp * int(world))(*[int(p) for p in peer_a])
        b = (_vp * int(world))(*[int(p) for p in peer_b])
        self._check(self._lib.ocl_sc_set_peer_xchg(self._h, int(rank), int(world), a, b), "ocl_sc_set_peer_xchg")

    def set_multicast_rho(self, local_ptr, multicast_ptr):
        self._check(self._lib.ocl_sc_set_multicast_rho(self._h, int(local_ptr), int(multicast_ptr)),
                    "ocl_sc_set_multicast_rho")

    def set_multicast_phi(self, local_ptr, multicast_ptr):
        self._check(self._lib.ocl_sc_set_multicast_phi(self._h, int(local_ptr), int(multicast_ptr)),
                    "ocl_sc_set_multicast_phi")

    def nvls_reduce_rho(self, stream=None):
        self._check(self._lib.ocl_sc_nvls_reduce_rho(self._h, _stream_ptr(stream)), "ocl_sc_nvls_reduce_rho")

    def mailbox_status(self, synchronise=True) -> int:
        """0: every fused exchange so far completed; 1 / 2 / 3: a momentum / extent / barrier exchange timed out."""
        st = C.c_int(0)
        self._check(self._lib.ocl_sc_mailbox_status(self._h, 1 if synchronise else 0, C.byref(st)),
                    "ocl_sc_mailbox_status")
        return int(st.value)

    def mailbox_exchange(self, which, stream=None):
        self._check(self._lib.ocl_sc_mailbox_exchange(self._h, int(which), _stream_ptr(stream)),
                    "ocl_sc_mailbox_exchange")

    def use_device_params(self, on: bool):
        self._check(self._lib.ocl_sc_use_device_params(self._h, 1 if on else 0), "ocl_sc_use_device_params")

    def set_kick_params(self, E_GeV, dz, mesh_draws=None, stream=None):
        self._check(self._lib.ocl_sc_set_kick_params(self._h, float(E_GeV), float(dz), _draws(mesh_draws),
                                                     _stream_ptr(stream)), "ocl_sc_set_kick_params")

    def stage_momentum(self, r, E_GeV, stream=None):
        ptr, ld, n = self._dev_rows(r)
        self._check(self._lib.ocl_sc_stage_momentum(self._h, ptr, ld, n, float(E_GeV), _stream_ptr(stream)),
                    "ocl_sc_stage_momentum")

    def stage_extent(self, r, q, E_GeV, mesh_draws=None, stream=None):
        ptr, ld, n = self._dev_rows(r, q)
        self._check(self._lib.ocl_sc_stage_extent(self._h, ptr, ld, q.data_ptr(), n, float(E_GeV),
                                                  _draws(mesh_draws), _stream_ptr(stream)), "ocl_sc_stage_extent")

    def defer_finish(self, on: bool):
        """NCCL fallback of a sharded kick: the sweeps only reduce locally; the caller all-reduces the
        collective buffers and calls ``stage_finish``."""
        self._check(self._lib.ocl_sc_defer_finish(self._h, 1 if on else 0), "ocl_sc_defer_finish")

    def set_deterministic(self, on: bool):
        """Ordered deposit (np.bincount's summation order; bit-identical rho from run to run) on / off."""
        self._check(self._lib.ocl_sc_set_deterministic(self._h, 1 if on else 0), "ocl_sc_set_deterministic")

    def stage_finish(self, which, E_GeV, mesh_draws=None, stream=None):
        self._check(self._lib.ocl_sc_stage_finish(self._h, int(which), float(E_GeV), _draws(mesh_draws),
                                                  _stream_ptr(stream)), "ocl_sc_stage_finish")

    def stage_deposit(self, r, q, E_GeV, mesh_draws=None, stream=None):
        ptr, ld, n = self._dev_rows(r, q)
        self._check(self._lib.ocl_sc_stage_deposit(self._h, ptr, ld, q.data_ptr(), n, float(E_GeV),
                                                   _draws(mesh_draws), _stream_ptr(stream)), "ocl_sc_stage_deposit")

    def stage_solve(self, mesh_draws=None, stream=None):
        self._check(self._lib.ocl_sc_stage_solve(self._h, _draws(mesh_draws), _stream_ptr(stream)),
                    "ocl_sc_stage_solve")

    def stage_kick(self, r, E_GeV, dz, mesh_draws=None, stream=None):
        ptr, ld, n = self._dev_rows(r)
        self._check(self._lib.ocl_sc_stage_kick(self._h, ptr, ld, n, float(E_GeV), float(dz), _draws(mesh_draws),
                                                _stream_ptr(stream)), "ocl_sc_stage_kick")

    # -- taps -----------------------------------------------------------------
    def geometry(self) -> dict:
        out = (C.c_double * 24)()
        self._check(self._lib.ocl_sc_get_geometry(self._h, out), "ocl_sc_get_geometry")
        a = np.array(out[:])
        return dict(T=a[:9].reshape(3, 3), pav=a[9], gamma0=a[10], beta0=a[11], steps=a[12:15], X_off=a[15:18],
                    sum_q=a[18], count=a[19])

    def _grid(self, fn, what):
        out = np.empty(self.nmesh, dtype=np.float64)
        self._check(fn(self._h, out.ctypes.data), what)
        return out

    def rho(self):
        return self._grid(self._lib.ocl_sc_get_rho, "ocl_sc_get_rho")

    def phi(self):
        return self._grid(self._lib.ocl_sc_get_phi, "ocl_sc_get_phi")

    def green(self):
        return self._grid(self._lib.ocl_sc_get_green, "ocl_sc_get_green")

    def field_at_particles(self, r, q, E_GeV, mesh_draws=None, stream=None):
        import torch
        ptr, ld, n = self._dev_rows(r, q)
        out = torch.empty((n, 3), dtype=torch.float64, device=r.device)
        self._check(self._lib.ocl_sc_field_at_particles(self._h, ptr, ld, q.data_ptr(), n, float(E_GeV),
                                                        _draws(mesh_draws), out.data_ptr(), _stream_ptr(stream)),
                    "ocl_sc_field_at_particles")
        return out

    def mad_to_cartesian(self, r, E_GeV, stream=None):
        import torch
        ptr, ld, n = self._dev_rows(r)
        xp = torch.empty((6, n), dtype=torch.float64, device=r.device)
        self._check(self._lib.ocl_sc_mad_to_cartesian(self._h, ptr, ld, n, float(E_GeV), xp.data_ptr(), n,
                                                      _stream_ptr(stream)), "ocl_sc_mad_to_cartesian")
        return xp

    def cartesian_to_mad(self, xp, E_GeV, stream=None):
        import torch
        ptr, ld, n = self._dev_rows(xp)
        r = torch.empty((6, n), dtype=torch.float64, device=xp.device)
        self._check(self._lib.ocl_sc_cartesian_to_mad(self._h, ptr, ld, n, float(E_GeV), r.data_ptr(), n,
                                                      _stream_ptr(stream)), "ocl_sc_cartesian_to_mad")
        return r

    def potential_host(self, rho: np.ndarray, steps) -> np.ndarray:
        rho = np.ascontiguousarray(rho, dtype=np.float64)
        if rho.shape != self.nmesh:
            raise ValueError(f"rho must have shape {self.nmesh}")
        st = (C.c_double * 3)(*[float(s) for s in steps])
        out = np.empty(self.nmesh, dtype=np.float64)
        self._check(self._lib.ocl_sc_potential_host(self._h, rho.ctypes.data, st, out.ctypes.data),
                    "ocl_sc_potential_host")
        return out

    # -- neighbours of the kick in the tracking loop -----------------------------
    def map_apply(self, r, R, B=None, T=None, stream=None):
        """rparticles <- R r + T:rr + B in place (first- or second-order transfer map)."""
        ptr, ld, n = self._dev_rows(r)
        Rc = np.ascontiguousarray(R, dtype=np.float64).reshape(36)
        Bc = None if B is None else np.ascontiguousarray(B, dtype=np.float64).reshape(6)
        Tc = None if T is None else np.ascontiguousarray(T, dtype=np.float64).reshape(216)
        as_p = lambda a: None if a is None else a.ctypes.data_as(_dp)
        self._check(self._lib.ocl_sc_map_apply(self._h, ptr, ld, n, as_p(Rc), as_p(Bc), as_p(Tc),
                                               _stream_ptr(stream)), "ocl_sc_map_apply")

    def cavity_apply(self, r, R, B, coef, mode=1, stream=None):
        """RF cavity body: R r + B, then the longitudinal RF map with the 7 host-computed scalars."""
        ptr, ld, n = self._dev_rows(r)
        Rc = np.ascontiguousarray(R, dtype=np.float64).reshape(36)
        Bc = np.zeros(6) if B is None else np.ascontiguousarray(B, dtype=np.float64).reshape(6)
        cc = np.ascontiguousarray(coef, dtype=np.float64).reshape(7)
        self._check(self._lib.ocl_sc_cavity_apply(self._h, ptr, ld, n, Rc.ctypes.data_as(_dp), Bc.ctypes.data_as(_dp),
                                                  cc.ctypes.data_as(_dp), int(mode), _stream_ptr(stream)),
                    "ocl_sc_cavity_apply")

    def aperture_cut(self, r, q, ids, kind, row, params, r_out, q_out, ids_out, lost_out, stream=None) -> int:
        """Ordered compaction of the particles inside the aperture into the *_out tensors; returns the survivor count."""
        ptr, ld, n = self._dev_rows(r, q)
        prm = (C.c_double * 4)(*[float(v) for v in params])
        n_out = _ll(0)
        self._check(self._lib.ocl_sc_aperture_cut(self._h, ptr, ld, q.data_ptr(), ids.data_ptr(), n, int(kind), int(row), prm,
                                                  r_out.data_ptr(), r_out.stride(0), q_out.data_ptr(), ids_out.data_ptr(),
                                                  lost_out.data_ptr(), C.byref(n_out), _stream_ptr(stream)),
                    "ocl_sc_aperture_cut")
        return int(n_out.value)

    MOMENT_KEYS = ("x", "px", "y", "py", "tau", "p", "xx", "xpx", "pxpx", "yy", "ypy", "pypy", "tautau", "pp",
                   "xy", "pxpy", "xpy", "ypx")

    def beam_moments(self, r, stream=None) -> dict:
        ptr, ld, n = self._dev_rows(r)
        out = (C.c_double * 18)()
        self._check(self._lib.ocl_sc_beam_moments(self._h, ptr, ld, n, out, _stream_ptr(stream)),
                    "ocl_sc_beam_moments")
        return dict(zip(self.MOMENT_KEYS, out[:]))

    def beam_moments_device(self, r, q, out, stream=None):
        """The same two passes, result left on the device: ``out`` is a contiguous CUDA float64 tensor of at least
        19 elements (18 moments in MOMENT_KEYS order, then sum q); no host synchronisation."""
        ptr, ld, n = self._dev_rows(r, q)
        if not (out.is_cuda and out.dtype == r.dtype and out.is_contiguous() and out.numel() >= 19):
            raise TypeError("out must be a contiguous CUDA float64 tensor with at least 19 elements")
        self._check(self._lib.ocl_sc_beam_moments_device(self._h, ptr, ld, n, q.data_ptr(), out.data_ptr(),
                                                         _stream_ptr(stream)), "ocl_sc_beam_moments_device")

    # -- longitudinal space charge (sc.py:261-599) ----------------------------------
    LSC_STAT_KEYS = ("n", "mean_tau", "m2_tau", "min_tau", "max_tau", "sum_q", "sum_x", "sum_y")
    LSC_PARAM_KEYS = ("slice_min", "slice_max", "x_shift", "y_shift", "a", "ds", "nb", "sigma_s", "K", "q", "v",
                      "gamma", "dz", "und", "pc_ref", "step_profile", "n_total")

    def lsc_stats(self, r, q, stream=None) -> dict:
        """Sweep A of an LSC kick (synchronous): this buffer's n, mean and centred square sum of tau,
        min/max tau, sum q, sum x, sum y."""
        ptr, ld, n = self._dev_rows(r, q)
        out = (C.c_double * 8)()
        self._check(self._lib.ocl_sc_lsc_stats(self._h, ptr, ld, n, q.data_ptr(), out, _stream_ptr(stream)),
                    "ocl_sc_lsc_stats")
        return dict(zip(self.LSC_STAT_KEYS, out[:]))

    def _lsc_params(self, params: dict):
        return (C.c_double * 17)(*[float(params[k]) for k in self.LSC_PARAM_KEYS])

    def lsc_kick(self, r, params: dict, stream=None):
        ptr, ld, n = self._dev_rows(r)
        self._check(self._lib.ocl_sc_lsc_kick(self._h, ptr, ld, n, self._lsc_params(params), _stream_ptr(stream)),
                    "ocl_sc_lsc_kick")

    def lsc_deposit(self, r, params: dict, stream=None):
        ptr, ld, n = self._dev_rows(r)
        self._check(self._lib.ocl_sc_lsc_deposit(self._h, ptr, ld, n, self._lsc_params(params),
                                                 _stream_ptr(stream)), "ocl_sc_lsc_deposit")

    def lsc_solve_kick(self, r, params: dict, stream=None):
        ptr, ld, n = self._dev_rows(r)
        self._check(self._lib.ocl_sc_lsc_solve_kick(self._h, ptr, ld, n, self._lsc_params(params),
                                                    _stream_ptr(stream)), "ocl_sc_lsc_solve_kick")

    def lsc_kick_async(self, r, q, gamma, v, pc_ref, dz, und, bounds, smooth_param, step_profile, n_total=0,
                       stream=None):
        """One LSC kick with the grid derived on the device: no host synchronisation."""
        ptr, ld, n = self._dev_rows(r, q)
        hp = (C.c_double * 10)(float(gamma), float(v), float(pc_ref), float(dz), float(und), float(bounds[0]),
                               float(bounds[1]), float(smooth_param), 1.0 if step_profile else 0.0, float(n_total))
        self._check(self._lib.ocl_sc_lsc_kick_async(self._h, ptr, ld, n, q.data_ptr(), hp, _stream_ptr(stream)),
                    "ocl_sc_lsc_kick_async")

    def lsc_async_status(self, synchronise=True) -> int:
        """0: every asynchronous LSC kick so far was applied; 1 / 2: one was skipped on the device (grid beyond
        the asynchronous capacity / packed word too narrow) and left the particles untouched.  Clears the flag."""
        st = C.c_int(0)
        self._check(self._lib.ocl_sc_lsc_async_status(self._h, 1 if synchronise else 0, C.byref(st)),
                    "ocl_sc_lsc_async_status")
        return int(st.value)

    def lsc_last_params(self) -> dict:
        """Scalars the device derived for the last asynchronous kick (synchronises)."""
        out = (C.c_double * 17)()
        self._check(self._lib.ocl_sc_lsc_last_params(self._h, out), "ocl_sc_lsc_last_params")
        d = dict(zip(self.LSC_PARAM_KEYS, out[:]))
        d["nb"], d["K"] = int(d["nb"]), int(d["K"])
        return d

    def lsc_profile(self, nb: int) -> dict:
        cur, wake = np.empty(nb), np.empty(nb)
        sig = C.c_double()
        self._check(self._lib.ocl_sc_lsc_get_profile(self._h, int(nb), cur.ctypes.data_as(_dp),
                                                     wake.ctypes.data_as(_dp), C.byref(sig)),
                    "ocl_sc_lsc_get_profile")
        return dict(current=cur, W=wake, sigma=sig.value)

    # -- timers ---------------------------------------------------------------
    def enable_timers(self, on=True):
        self._check(self._lib.ocl_sc_enable_timers(self._h, 1 if on else 0), "ocl_sc_enable_timers")

    def timers(self) -> dict:
        out = (C.c_double * 8)()
        self._check(self._lib.ocl_sc_get_timers(self._h, out), "ocl_sc_get_timers")
        keys = ("momentum", "extent", "deposit", "solve", "field", "kick", "total")
        return {k: out[i] for i, k in enumerate(keys)}

    def launch_count(self) -> int:
        return int(self._lib.ocl_sc_launch_count(self._h))


def _wrap_device_doubles(ptr: int, count: int, device: int):
    """Zero-copy torch tensor over ``count`` doubles at device address ``ptr``."""
    import torch

    class _Span:
        pass

    span = _Span()
    span.__cuda_array_interface__ = {"shape": (int(count),), "typestr": "<f8", "data": (int(ptr), False),
                                     "version": 2}
    return torch.as_tensor(span, device=torch.device("cuda", device))
