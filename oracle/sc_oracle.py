"""CPU oracle for the 3D space-charge kick -- TEST INFRASTRUCTURE ONLY.

This module is a numpy restatement of the algorithm in the reference's
``ocelot/cpbd/sc.py`` (``SpaceCharge``) and ``ocelot/cpbd/coord_transform.py``.
It exists to check the CUDA path; nothing under ``ocelot_b200/`` may import it.
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py`` (the
``cpu_baseline`` leg and ``--impl reference``) may use it, and only as the
checker / the CPU timing arm.

Parity status: PINNED.  ``oracle/make_golden.py`` (run in the build container,
where ``/root/reference`` is importable) executes the unmodified reference on
seeded inputs and stores its stage outputs under ``tests/golden/``;
``tests/test_oracle_golden.py`` checks every function below against them.

Every function cites the reference lines it restates (paths relative to the
reference checkout).  Arithmetic is kept in the reference's operation order so
that ``fft="reference"`` reproduces the reference bit for bit on the same
numpy/scipy; ``fft="padded"`` swaps the (2n-1)^3 complex transform for a
power-of-two real transform, which is the same linear convolution.
"""
from __future__ import annotations

import numpy as np

try:  # scipy is only needed for the multi-threaded padded FFT
    import scipy.fft as _sfft
except Exception:  # pragma: no cover
    _sfft = None

# ---------------------------------------------------------------------------
# constants: ocelot/common/globals.py:13-24 (same expressions => same bits)
# ---------------------------------------------------------------------------
PI = 3.141592653589793
C_LIGHT = 299792458.0
Q_E = 1.6021766208e-19
M_E_KG = 9.10938215e-31
M_E_EV = M_E_KG * C_LIGHT ** 2 / Q_E
M_E_GEV = M_E_EV / 1e+9
MU_0 = 4 * PI * 1e-7
EPS_0 = 1 / MU_0 / C_LIGHT ** 2


# ---------------------------------------------------------------------------
# coordinate transforms
# ---------------------------------------------------------------------------
def mad_to_cartesian(r: np.ndarray, gamref: float) -> np.ndarray:
    """(x, x', y, y', tau, delta) -> (x, y, z, px, py, pz[eV/c]).

    Restates ``xxstg_2_xp_mad`` (coord_transform.py:57-96, numpy branch
    :68-70, :73-81, :90-95).  ``r`` is (6, N); returns a new (6, N) array.
    """
    n = r.shape[1]
    betaref = np.sqrt(1 - gamref ** -2)
    gam = (betaref * r[5] + 1) * gamref                       # :68
    bet = np.sqrt(1 - gam ** -2)                              # :69
    pz_rel = np.sqrt(((gam * bet) / (gamref * betaref)) ** 2 - r[1] ** 2 - r[3] ** 2)  # :70
    dirs = np.c_[r[1] / pz_rel, r[3] / pz_rel, np.ones(n)]    # :72
    dirs = dirs / np.linalg.norm(dirs, 2, 1).reshape((n, 1))  # :74, :77
    u0, u1, u2 = dirs[:, 0], dirs[:, 1], dirs[:, 2]
    out = np.zeros((6, n))
    out[0] = r[0] - u0 * bet * r[4]                           # :90
    out[1] = r[2] - u1 * bet * r[4]                           # :91
    out[2] = -u2 * bet * r[4]                                 # :92
    out[3] = u0 * gam * bet * M_E_EV                          # :93
    out[4] = u1 * gam * bet * M_E_EV                          # :94
    out[5] = u2 * gam * bet * M_E_EV                          # :95
    return out


def cartesian_to_mad(xp: np.ndarray, r_out: np.ndarray, gamref: float) -> np.ndarray:
    """(x, y, z, px, py, pz) -> MAD rows, written in place into ``r_out``.

    Restates ``xp_2_xxstg_mad`` (coord_transform.py:16-54, numpy branch
    :27-28, :31, :34, :47-53).
    """
    n = xp.shape[1]
    pref = M_E_EV * np.sqrt(gamref ** 2 - 1)                  # :19
    betaref = np.sqrt(1 - gamref ** -2)                       # :20
    mom = np.c_[xp[3], xp[4], xp[5]]                          # :21
    gam = np.sqrt(1 + np.sum(mom * mom, 1) / M_E_EV ** 2)     # :27
    bet = np.sqrt(1 - gam ** -2)                              # :28
    mom = mom / np.linalg.norm(mom, 2, 1).reshape((n, 1))     # :31, :34
    u0, u1, u2 = mom[:, 0], mom[:, 1], mom[:, 2]
    cdt = -xp[2] / (bet * u2)                                 # :47
    r_out[0] = xp[0] + bet * u0 * cdt                         # :48
    r_out[2] = xp[1] + bet * u1 * cdt                         # :49
    r_out[5] = (gam / gamref - 1) / betaref                   # :50
    r_out[4] = cdt                                            # :51
    r_out[1] = xp[3] / pref                                   # :52
    r_out[3] = xp[4] / pref                                   # :53
    return r_out


# ---------------------------------------------------------------------------
# bunch frame
# ---------------------------------------------------------------------------
def bunch_frame(mom: np.ndarray):
    """Mean-momentum frame.  Restates sc.py:224-231 and :237-239.

    ``mom`` is the (3, N) block of Cartesian momenta.  Returns
    ``(T, pav, gamma0, beta0)``; the columns of ``T`` are t1, t2, t3.
    """
    t3 = np.mean(mom, axis=1)
    pav = np.linalg.norm(t3)
    t3 = t3 / pav
    t1 = np.cross(np.array([0, 1, 0]), t3)
    t1 = t1 / np.linalg.norm(t1)
    t2 = np.cross(t3, t1)
    T = np.c_[t1, t2, t3]
    gamma0 = np.sqrt((pav / M_E_EV) ** 2 + 1)
    beta0 = np.sqrt(1 - gamma0 ** -2)
    return T, pav, gamma0, beta0


# ---------------------------------------------------------------------------
# integrated Green's function
# ---------------------------------------------------------------------------
def igf_kernel(nxyz, steps) -> np.ndarray:
    """Integrated Green's function on the n^3 block of non-negative offsets.

    Restates ``SpaceCharge.sym_kernel`` (sc.py:109-133): antiderivative at the
    half-offset points ``h*i - h/2`` (i = 0..n), term order of sc.py:124-126,
    then the 8-corner alternating difference in the order of sc.py:128-131.
    """
    nx, ny, nz = int(nxyz[0]), int(nxyz[1]), int(nxyz[2])
    hx, hy, hz = steps[0], steps[1], steps[2]
    x = hx * np.r_[0:nx + 1] - hx / 2
    y = hy * np.r_[0:ny + 1] - hy / 2
    z = hz * np.r_[0:nz + 1] - hz / 2
    x, y, z = np.ix_(x, y, z)
    r = np.sqrt(x * x + y * y + z * z)
    G = -x * x * 0.5 * np.arctan(y * z / (x * r))
    G = G + y * z * np.log(x + r)
    G = G - y * y * 0.5 * np.arctan(z * x / (y * r))
    G = G + z * x * np.log(y + r)
    G = G - z * z * 0.5 * np.arctan(x * y / (z * r))
    G = G + x * y * np.log(z + r)
    hi_x, lo_x = slice(1, nx + 1), slice(0, nx)
    hi_y, lo_y = slice(1, ny + 1), slice(0, ny)
    hi_z, lo_z = slice(1, nz + 1), slice(0, nz)
    K = G[hi_x, hi_y, hi_z] - G[lo_x, hi_y, hi_z]
    K = K - G[hi_x, lo_y, hi_z]
    K = K + G[lo_x, lo_y, hi_z]
    K = K - G[hi_x, hi_y, lo_z]
    K = K + G[lo_x, hi_y, lo_z]
    K = K + G[hi_x, lo_y, lo_z]
    K = K - G[lo_x, lo_y, lo_z]
    return K


def fft_size(n: int) -> int:
    """Smallest power of two >= 2n-1 (the device grid size M)."""
    m = 1
    while m < 2 * n - 1:
        m *= 2
    return m


def mirrored_kernel(K1: np.ndarray, shape) -> np.ndarray:
    """Place K1[|d|] at d mod M for every axis (sc.py:145-149 generalised to
    any period M >= 2n-1; entries with n <= d <= M-n stay zero)."""
    nx, ny, nz = K1.shape
    K2 = np.zeros(shape)
    K2[:nx, :ny, :nz] = K1
    K2[:nx, :ny, shape[2] - nz + 1:] = K2[:nx, :ny, nz - 1:0:-1]
    K2[:nx, shape[1] - ny + 1:, :] = K2[:nx, ny - 1:0:-1, :]
    K2[shape[0] - nx + 1:, :, :] = K2[nx - 1:0:-1, :, :]
    return K2


def poisson_potential(rho: np.ndarray, steps, fft: str = "reference", workers: int = 1) -> np.ndarray:
    """Open-boundary potential of ``rho`` by Hockney convolution with the IGF.

    Restates ``SpaceCharge.potential`` (sc.py:135-168).  ``fft="reference"``
    uses numpy's complex transform on the (2n-1)^3 grid exactly as sc.py:164;
    ``fft="padded"`` uses a real transform on the power-of-two grid (same
    linear convolution, round-off level difference).
    """
    nx, ny, nz = rho.shape
    hx, hy, hz = steps[0], steps[1], steps[2]
    K1 = igf_kernel(rho.shape, steps)
    if fft == "reference":
        shape = (2 * nx - 1, 2 * ny - 1, 2 * nz - 1)
        pad = np.zeros(shape)
        pad[:nx, :ny, :nz] = rho
        K2 = mirrored_kernel(K1, shape)
        conv = np.real(np.fft.ifftn(np.fft.fftn(pad) * np.fft.fftn(K2)))
    elif fft == "padded":
        shape = (fft_size(nx), fft_size(ny), fft_size(nz))
        pad = np.zeros(shape)
        pad[:nx, :ny, :nz] = rho
        K2 = mirrored_kernel(K1, shape)
        conv = _sfft.irfftn(_sfft.rfftn(pad, workers=workers) * _sfft.rfftn(K2, workers=workers),
                            s=shape, workers=workers)
    else:
        raise ValueError(fft)
    return conv[:nx, :ny, :nz] / (4 * PI * EPS_0 * hx * hy * hz)


# ---------------------------------------------------------------------------
# mesh geometry, deposit, field, gather
# ---------------------------------------------------------------------------
def mesh_geometry(X: np.ndarray, Q: np.ndarray, nxyz, mesh_scale=None, mesh_shift=None):
    """Cell sizes and the mesh origin.  Restates sc.py:173-186.

    ``X`` is (N, 3), already rotated and gamma-stretched.  ``mesh_scale`` /
    ``mesh_shift`` are the two ``random_mesh`` draws (sc.py:175, :185) or None.
    Returns ``(steps, X_off, Xg)`` with ``Xg`` the particle positions in cell
    units relative to the mesh origin.
    """
    nxyz = np.asarray(nxyz)
    extent = np.max(X, axis=0) - np.min(X, axis=0)
    if mesh_scale is not None:
        extent = extent * mesh_scale
    steps = extent / (nxyz - 3)
    Xg = X / steps
    X_min = np.min(Xg, axis=0)
    X_mid = np.dot(Q, Xg) / np.sum(Q)
    X_off = np.floor(X_min - X_mid) + X_mid
    if mesh_shift is not None:
        X_off = X_off + mesh_shift
    Xg = Xg - X_off
    return steps, X_off, Xg


def cell_index(Xg: np.ndarray, nxyz) -> np.ndarray:
    """Nearest-grid-point cell of each particle (x slowest, z fastest).
    Restates sc.py:191-192."""
    ny, nz = int(nxyz[1]), int(nxyz[2])
    cell = np.int_(np.floor(Xg) + 1)
    return np.int_(cell[:, 0] * (nz * ny) + cell[:, 1] * nz + cell[:, 2])


def deposit_ngp(idx: np.ndarray, Q: np.ndarray, nxyz) -> np.ndarray:
    """Nearest-grid-point charge deposit.  Restates sc.py:193."""
    nx, ny, nz = int(nxyz[0]), int(nxyz[1]), int(nxyz[2])
    return np.bincount(idx, Q, nx * ny * nz).reshape((nx, ny, nz))


def staggered_field(phi: np.ndarray, steps):
    """Backward differences of the potential, last plane zero.
    Restates sc.py:195-200."""
    nx, ny, nz = phi.shape
    Ex = np.zeros(phi.shape)
    Ey = np.zeros(phi.shape)
    Ez = np.zeros(phi.shape)
    Ex[:nx - 1, :, :] = (phi[:nx - 1, :, :] - phi[1:nx, :, :]) / steps[0]
    Ey[:, :ny - 1, :] = (phi[:, :ny - 1, :] - phi[:, 1:ny, :]) / steps[1]
    Ez[:, :, :nz - 1] = (phi[:, :, :nz - 1] - phi[:, :, 1:nz]) / steps[2]
    return Ex, Ey, Ez


def trilinear(F: np.ndarray, c0, c1, c2) -> np.ndarray:
    """Order-1 interpolation of grid ``F`` at fractional coordinates, returning
    0 for any coordinate outside [0, n-1].

    Restates what ``scipy.ndimage.map_coordinates(order=1)`` computes at
    sc.py:202-204 (mode='constant', cval=0): weights (1-t, t) per axis, each
    corner value multiplied by its x, y, z weights in that order, corners
    accumulated with z fastest (scipy's NI_GeometricTransform loop order).
    """
    nx, ny, nz = F.shape
    inside = ((c0 >= 0) & (c0 <= nx - 1) & (c1 >= 0) & (c1 <= ny - 1)
              & (c2 >= 0) & (c2 <= nz - 1))
    s0 = np.where(inside, c0, 0.0)
    s1 = np.where(inside, c1, 0.0)
    s2 = np.where(inside, c2, 0.0)
    i0 = np.floor(s0).astype(np.int64)
    i1 = np.floor(s1).astype(np.int64)
    i2 = np.floor(s2).astype(np.int64)
    t0 = s0 - i0
    t1 = s1 - i1
    t2 = s2 - i2
    j0 = np.minimum(i0 + 1, nx - 1)
    j1 = np.minimum(i1 + 1, ny - 1)
    j2 = np.minimum(i2 + 1, nz - 1)
    acc = np.zeros(c0.shape)
    for a, wa in ((i0, 1 - t0), (j0, t0)):
        for b, wb in ((i1, 1 - t1), (j1, t1)):
            for c, wc in ((i2, 1 - t2), (j2, t2)):
                acc = acc + F[a, b, c] * wa * wb * wc
    return np.where(inside, acc, 0.0)


def field_at_particles(X: np.ndarray, Q: np.ndarray, gamma0: float, nxyz,
                       mesh_scale=None, mesh_shift=None, fft: str = "reference",
                       workers: int = 1, taps: dict | None = None) -> np.ndarray:
    """Rest-frame field sampled at the particles, lab-frame scaled.

    Restates ``SpaceCharge.el_field`` (sc.py:170-205).  ``X`` (N, 3) holds the
    rotated lab positions and is NOT modified (the reference stretches its
    argument in place at sc.py:172).  If ``taps`` is a dict, stage outputs are
    stored in it (steps, X_off, idx, rho, phi).
    """
    nxyz = np.asarray(nxyz)
    X = np.array(X, dtype=np.float64, copy=True)
    X[:, 2] = X[:, 2] * gamma0                                 # :172
    steps, X_off, Xg = mesh_geometry(X, Q, nxyz, mesh_scale, mesh_shift)
    idx = cell_index(Xg, nxyz)
    rho = deposit_ngp(idx, Q, nxyz)
    phi = poisson_potential(rho, steps, fft=fft, workers=workers)
    Ex, Ey, Ez = staggered_field(phi, steps)
    E = np.zeros((X.shape[0], 3))
    E[:, 0] = trilinear(Ex, Xg[:, 0], Xg[:, 1] + 0.5, Xg[:, 2] + 0.5) * gamma0   # :202
    E[:, 1] = trilinear(Ey, Xg[:, 0] + 0.5, Xg[:, 1], Xg[:, 2] + 0.5) * gamma0   # :203
    E[:, 2] = trilinear(Ez, Xg[:, 0] + 0.5, Xg[:, 1] + 0.5, Xg[:, 2])            # :204
    if taps is not None:
        taps.update(steps=steps, X_off=X_off, idx=idx, rho=rho, phi=phi)
    return E


# ---------------------------------------------------------------------------
# the kick
# ---------------------------------------------------------------------------
def sc_kick(r: np.ndarray, q: np.ndarray, E_GeV: float, dz: float, nmesh_xyz,
            mesh_scale=None, mesh_shift=None, fft: str = "reference", workers: int = 1,
            taps: dict | None = None) -> None:
    """One space-charge kick, in place on ``r`` (6, N).

    Restates ``SpaceCharge.apply`` (sc.py:208-251).
    """
    if dz == 0:                                               # :210-212
        return
    nmesh = np.array(nmesh_xyz)
    gamref = E_GeV / M_E_GEV                                  # :214
    betref = np.sqrt(1 - gamref ** -2)                        # :215-216
    xp = mad_to_cartesian(r, gamref)                          # :221
    T, pav, gamma0, beta0 = bunch_frame(xp[3:6])              # :224-239
    xyz = np.dot(xp[0:3].T, T)                                # :233
    xp[3:6] = np.dot(xp[3:6].T, T).T                          # :234
    E = field_at_particles(xyz, q, gamma0, nmesh, mesh_scale, mesh_shift,
                           fft=fft, workers=workers, taps=taps)   # :241
    cdT = dz / betref                                         # :244
    xp[3] = xp[3] + cdT * (1 - beta0 * beta0) * E[:, 0]       # :246
    xp[4] = xp[4] + cdT * (1 - beta0 * beta0) * E[:, 1]       # :247
    xp[5] = xp[5] + cdT * E[:, 2]                             # :248
    xp[3:6] = np.dot(xp[3:6].T, np.transpose(T)).T            # :249-250
    if taps is not None:
        taps.update(T=T, pav=pav, gamma0=gamma0, beta0=beta0, Exyz=E)
    cartesian_to_mad(xp, r, gamref)                           # :251


# ---------------------------------------------------------------------------
# synthetic bunch (the BASELINE configs' input generator)
# ---------------------------------------------------------------------------
def _inverse_cdf(x: np.ndarray, y: np.ndarray):
    """Restates ``invert_cdf`` (ocelot/common/math_op.py) for the Gaussian
    profile used by ``generate_parray``: cumulative trapezoid -> normalise ->
    linear interpolation of x over the CDF."""
    from scipy import integrate, interpolate
    cum = integrate.cumulative_trapezoid(y, x, initial=0)
    cum = cum / cum[-1]
    return interpolate.interp1d(cum, x, bounds_error=False, fill_value=(x[0], x[-1]))


def gaussian_bunch(nparticles: int, energy: float = 0.13, charge: float = 5e-9,
                   sigma_x: float = 1e-4, sigma_px: float = 2e-5, sigma_y=None, sigma_py=None,
                   sigma_tau: float = 1e-3, sigma_p: float = 1e-4, chirp: float = 0.01):
    """Gaussian ParticleArray contents, drawing from numpy's GLOBAL RNG in the
    reference's order.  Restates the default branch of ``generate_parray``
    (ocelot/cpbd/beam/generator.py:107-116, :118-135, :138, :150-159, :170-172).
    Returns ``(rparticles (6,N), q_array (N,), E)``.
    """
    sigma_y = sigma_x if sigma_y is None else sigma_y
    sigma_py = sigma_px if sigma_py is None else sigma_py
    x = np.random.randn(nparticles) * sigma_x
    px = np.random.randn(nparticles) * sigma_px
    y = np.random.randn(nparticles) * sigma_y
    py = np.random.randn(nparticles) * sigma_py
    s = np.linspace(-5 * sigma_tau, 5 * sigma_tau, num=500)
    prof = np.exp(-s ** 2 / (2. * sigma_tau ** 2))
    tau = _inverse_cdf(s, prof)(np.random.rand(nparticles))
    dp = np.random.randn(nparticles) * sigma_p
    if sigma_tau != 0:
        dp += chirp * tau / sigma_tau
    r = np.zeros((6, nparticles))
    r[0], r[1], r[2], r[3], r[4], r[5] = x, px, y, py, tau, dp
    q = np.ones(nparticles) * charge / nparticles
    return r, q, energy


# ---------------------------------------------------------------------------
# beam moments (the quantities the north-star's moment-level parity is stated on)
# ---------------------------------------------------------------------------
MOMENT_KEYS = ("x", "px", "y", "py", "tau", "p", "xx", "xpx", "pxpx", "yy", "ypy", "pypy",
               "tautau", "pp", "xy", "pxpy", "xpy", "ypx", "emit_x", "emit_y")


def beam_moments(r: np.ndarray) -> dict:
    """First and second moments and rms emittances of (6, N) MAD particles.

    Restates the default path (no bounds, no dispersion correction) of
    ``get_envelope`` (ocelot/cpbd/beam/analysis.py:72-76, :121-123, :125-166,
    :179-180)."""
    x, px, y, py, tau, p = r[0], r[1], r[2], r[3], r[4], r[5]
    m = {"p": np.mean(p)}
    factor = 1. - p - 0.5 * p * p + 0.5 * px * px + 0.5 * py * py      # :121
    px = px * factor
    py = py * factor
    m["x"], m["y"], m["px"], m["py"], m["tau"] = np.mean(x), np.mean(y), np.mean(px), np.mean(py), np.mean(tau)
    m["xx"] = np.mean((x - m["x"]) ** 2)
    m["xpx"] = np.mean((x - m["x"]) * (px - m["px"]))
    m["pxpx"] = np.mean((px - m["px"]) ** 2)
    m["yy"] = np.mean((y - m["y"]) ** 2)
    m["ypy"] = np.mean((y - m["y"]) * (py - m["py"]))
    m["pypy"] = np.mean((py - m["py"]) ** 2)
    m["tautau"] = np.mean((tau - m["tau"]) * (tau - m["tau"]))
    m["xy"] = np.mean((x - m["x"]) * (y - m["y"]))
    m["pxpy"] = np.mean((px - m["px"]) * (py - m["py"]))
    m["xpy"] = np.mean((x - m["x"]) * (py - m["py"]))
    m["ypx"] = np.mean((y - m["y"]) * (px - m["px"]))
    m["pp"] = np.mean((p - m["p"]) ** 2)
    m["emit_x"] = np.sqrt(m["xx"] * m["pxpx"] - m["xpx"] ** 2)
    m["emit_y"] = np.sqrt(m["yy"] * m["pypy"] - m["ypy"] ** 2)
    return {k: float(m[k]) for k in MOMENT_KEYS}


def replay_track(r: np.ndarray, q: np.ndarray, E_GeV: float, R, B, map_step, kick_dz, nmesh_xyz, kick,
                 after_step=None, T=None) -> None:
    """Replay a recorded first-order tracking run in place: for every step apply
    its transfer maps ``r <- R r + B`` (transfer_map.py:51-52) and then
    ``kick(r, q, E, dz, nmesh)`` -- the loop body of ``track()``
    (ocelot/cpbd/track.py:470-477) with the maps taken from a golden fixture."""
    map_step = np.asarray(map_step)
    for step, dz in enumerate(kick_dz):
        for m in np.nonzero(map_step == step)[0]:
            if T is None:
                r[:] = np.add(np.dot(R[m], r), B[m].reshape(6, 1))
            else:   # SecondTM.t_apply (second_order.py:31-39) with SecondOrderMult.numpy_apply (tm_utils.py:54-55)
                r[:] = np.matmul(R[m], r) + np.einsum('ijk,j...,k...->i...', T[m], r, r)
                r[:] = np.add(r, B[m].reshape(6, 1))
        if dz != 0:
            kick(r, q, E_GeV, float(dz), nmesh_xyz)
        if after_step is not None:
            after_step(step, r)


# ---------------------------------------------------------------------------
# RF cavity body and generic recorded-map replay (rows f1/f3 of SURVEY.md 8f)
# ---------------------------------------------------------------------------
def cavity_map(r: np.ndarray, R, B, v, phi_deg, freq, E, delta_length, length) -> float:
    """In place: the MAIN map of an RF cavity.  Restates ``CavityTM.map4cav``
    (ocelot/cpbd/transformations/cavity.py:29-128).  Returns the energy gain V cos(phi)."""
    if delta_length is not None:
        V = v * delta_length / length if length != 0 else v
        z = delta_length
    else:
        V, z = v, length
    beta0, igamma2, g0 = 1.0, 0.0, 1e10
    if E != 0.0:
        g0 = E / M_E_GEV
        igamma2 = 1.0 / (g0 * g0)
        beta0 = np.sqrt(1.0 - igamma2)
    phi = phi_deg * np.pi / 180.0
    X4, X5 = np.copy(r[4]), np.copy(r[5])
    r[:] = np.add(np.dot(R, r), np.asarray(B).reshape(6, 1))
    delta_e = V * np.cos(phi)
    E1 = E + delta_e
    T566, T556, T555 = 1.5 * z * igamma2 / (beta0 ** 3), 0.0, 0.0
    if E1 <= 0.0:
        r[4] += T566 * X5 * X5
        return delta_e
    k = 2.0 * np.pi * freq / C_LIGHT
    g1 = E1 / M_E_GEV
    beta1 = np.sqrt(1.0 - 1.0 / (g1 * g1))
    r[5] = (X5 * E * beta0 / (E1 * beta1)
            + V * beta0 / (E1 * beta1) * (np.cos(-X4 * beta0 * k + phi) - np.cos(phi)))
    dgamma = V / M_E_GEV
    dg = g1 - g0
    if abs(dg) < 1e-8 * abs(g0):
        if abs(np.cos(phi)) < 1e-3:
            T556 = 1.5 * z * k * dgamma / (beta0 ** 3 * g0 ** 3)
            T555 = 0.5 * z * k * k * dgamma * dgamma / (beta0 ** 3 * g0 ** 4)
    else:
        T566 = (z * (beta0 ** 3 * g0 ** 3 - beta1 ** 3 * g1 ** 3)
                / (2.0 * beta0 * beta1 ** 3 * g0 * (g0 - g1) * g1 ** 3))
        T556 = (beta0 * k * z * dgamma * g0 * (beta1 ** 3 * g1 ** 3 + beta0 * (g0 - g1 ** 3)) * np.sin(phi) /
                (beta1 ** 3 * g1 ** 3 * (g0 - g1) ** 2))
        T555 = (beta0 ** 2 * k ** 2 * z * dgamma / 2.0
                * (dgamma * (2.0 * g0 * g1 ** 3 * (beta0 * beta1 ** 3 - 1.0)
                             + g0 ** 2 + 3.0 * g1 ** 2 - 2.0) / (beta1 ** 3 * g1 ** 3 * (g0 - g1) ** 3) * np.sin(phi) ** 2
                   - (g1 * g0 * (beta1 * beta0 - 1.0) + 1.0)
                   / (beta1 * g1 * (g0 - g1) ** 2)
                   * np.cos(phi)))
    r[4] += T566 * X5 * X5 + T556 * X4 * X5 + T555 * X4 * X4
    return delta_e


def replay_recorded_maps(r, q, E, g, kick) -> float:
    """Replay a fixture written by oracle/make_golden.py::golden_injector_track: per step the
    recorded maps (kind 0 first order, 1 second order, 2 cavity body), then ``kick``.
    Returns the final beam energy."""
    map_step = np.asarray(g["map_step"])
    for step, dz in enumerate(g["kick_dz"]):
        for m in np.nonzero(map_step == step)[0]:
            kind = int(g["kind"][m])
            R, B = g["R"][m], g["B"][m]
            if kind == 2:
                v, phi, freq, dlen, length = g["cav"][m]
                cavity_map(r, R, B, v, phi, freq, E, None if np.isnan(dlen) else dlen, length)
            elif kind == 1:
                T = g["T"][int(g["Tidx"][m])]
                r[:] = np.matmul(R, r) + np.einsum('ijk,j...,k...->i...', T, r, r)
                r[:] = np.add(r, B.reshape(6, 1))
            else:
                r[:] = np.add(np.dot(R, r), B.reshape(6, 1))
            E += float(g["delta_e"][m])
        if dz != 0:
            kick(r, q, E, float(dz), g["nmesh"])
    return E
