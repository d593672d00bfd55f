"""CPU oracle for the longitudinal space-charge (LSC) kick -- TEST INFRASTRUCTURE ONLY.

numpy/scipy restatement of ``ocelot/cpbd/sc.py:261-599`` (class ``LSC``) and of the
helpers it calls: ``s_to_cur`` / ``s2cur_auxil`` (``ocelot/cpbd/beam/analysis.py:254-340``),
``convmode`` (``ocelot/cpbd/beam/beam_utils.py:59-68``) and ``signal_to_spectrum``
(``analysis.py:343-383``).  Nothing under ``ocelot_b200/`` may import it; only ``tests/``
use it, as the checker of the CUDA path (``ocelot_b200/csrc/sc_lsc.cu``).

Parity status: PINNED.  ``oracle/make_golden_lsc.py`` executes the unmodified reference
``LSC.apply`` on seeded bunches (Gaussian and step-profile models, with and without an
undulator section, plus kicks recorded from the reference's own LSC test lattice) and
stores inputs, stage outputs and kicked bunches in ``tests/golden/lsc_*.npz``;
``tests/test_oracle_lsc.py`` checks every function below against them.
"""
from __future__ import annotations

import numpy as np
from scipy.special import exp1, k1, factorial

PI = 3.141592653589793
C_LIGHT = 299792458.0
Q_E = 1.6021766208e-19
M_E_KG = 9.10938215e-31
M_E_EV = M_E_KG * C_LIGHT ** 2 / Q_E
M_E_GEV = M_E_EV / 1e+9
MU_0 = 4 * PI * 1e-7
EPS_0 = 1 / MU_0 / C_LIGHT ** 2
Z0 = 1. / (C_LIGHT * EPS_0)                      # globals.py:36


def cic_counts(A, a, ds, nbins):
    """s2cur_auxil (analysis.py:254-260) vectorised: unweighted first-order deposit."""
    cA = (A - a) / ds                             # analysis.py:324
    I = np.int_(np.floor(cA))                     # :325
    xiA = 1 + I - cA                              # :326
    I = np.minimum(I, nbins - 1)                  # :257-258
    C = np.bincount(I, xiA, nbins + 1) + np.bincount(I + 1, 1 - xiA, nbins + 1)
    return C[:nbins]


def current_grid(tau_min, tau_max, sigma):
    """Grid definition of s_to_cur for ds=None, N=None (analysis.py:293-321).
    Returns (a, ds, nbins) with nbins = N + 1 grid points x_j = j*ds + a."""
    Nsigma = 3
    a = tau_min
    b = tau_max
    if sigma is not None:
        a -= Nsigma * sigma
        b += Nsigma * sigma
    if sigma is not None and sigma > 0:
        ds = 0.25 * sigma
    else:
        ds = (b - a) / 1000.0
    N = int(np.ceil((b - a) / ds))
    ds = (b - a) / N
    return a, ds, N + 1


def smoothing_taps(sigma, ds):
    """Gaussian taps of s_to_cur (analysis.py:330-333); None when smoothing is off."""
    if sigma is None or not sigma > 0:
        return None
    K = int(np.floor(3 * sigma / ds + 0.5))
    G = np.exp(-0.5 * (np.arange(-K, K + 1) * ds / sigma) ** 2)
    return G / np.sum(G)


def s_to_cur(A, sigma, q0, v):
    """analysis.py:266-340 with ds=None, N=None.  Returns B (nbins, 2): [s, I(s)]."""
    a, ds, nbins = current_grid(np.min(A), np.max(A), sigma)
    B = np.zeros((nbins, 2))
    B[:, 0] = np.arange(0, (nbins - 1 + 0.5) * ds, ds) + a        # :321
    C = cic_counts(A, a, ds, nbins)
    G = smoothing_taps(sigma, ds)
    if G is not None:
        i = int(np.floor(len(G) * 0.5))                            # beam_utils.py:63
        B[:, 1] = np.convolve(C, G)[i:nbins + i]                   # :66-67
    else:
        B[:, 1] = C
    koef = q0 * v / (ds * np.sum(B[:, 1]))                          # analysis.py:338
    B[:, 1] = koef * B[:, 1]
    return B


def imp_lsc(gamma, sigma, w, dz):
    """Round Gaussian beam impedance (sc.py:299-340)."""
    eps = 1e-16
    ass = 40.0
    alpha = w * sigma / (gamma * C_LIGHT)
    alpha2 = alpha * alpha
    inda = np.where(alpha2 > ass)[0]
    ind = np.where((alpha2 <= ass) & (alpha2 >= eps))[0]
    T = np.zeros(w.shape)
    T[ind] = np.exp(alpha2[ind]) * exp1(alpha2[ind])
    x = alpha2[inda]
    k = 0
    for i in range(10):
        k += (-1) ** i * factorial(i) / (x ** (i + 1))
    T[inda] = k
    return 1j * Z0 / (4 * PI * C_LIGHT * gamma ** 2) * w * T * dz


def imp_step_lsc(gamma, rb, w, dz):
    """Uniform (step-profile) beam impedance (sc.py:342-369); ``w`` is modified in place
    like in the reference."""
    indx = np.where(w < 1e-7)[0]
    w[indx] = 1e-7
    x = w * rb / (C_LIGHT * gamma)
    Z = 1j * Z0 * C_LIGHT / (4 * w * rb * rb) * dz * (1 - x * k1(x))
    Z[indx] = 0
    return Z


def wake_lsc(s, bunch, gamma, sigma, dz, step_profile=False, K_max=0, fill_factor=0):
    """sc.py:418-474 (with signal_to_spectrum analysis.py:376-383 and impedance2wake
    sc.py:401-416 inlined)."""
    ds = s[1] - s[0]
    dt = ds / C_LIGHT
    nb = len(s)
    n = nb * 2
    f = 1 / dt * np.arange(0, n) / n
    imp = imp_step_lsc if step_profile else imp_lsc
    Za = imp(gamma, sigma, f[0:nb] * 2 * np.pi, dz) * (1 + 0.5 * K_max * K_max * fill_factor)
    bunch1 = np.append(bunch, np.zeros(nb))
    Zb = dt * np.fft.fft(bunch1 * C_LIGHT, n)
    Z = np.zeros(n, dtype=complex)
    Z[0:nb] = Za * Zb[0:nb]
    Z[nb:n] = np.flipud(np.conj(Z[0:nb]))
    df = f[1] - f[0]
    wa = n * df * np.fft.irfft(Z, n)
    return -wa[0:nb]


def lsc_stages(r, q_array, E_GeV, dz, step_profile=False, smooth_param=0.1, bounds=(-0.4, 0.4),
               K_max=0, fill_factor=0):
    """Everything LSC.apply computes before touching the particles (sc.py:566-592)."""
    tau = r[4]
    mean_tau = np.mean(tau)
    sigma_tau = np.std(tau)
    slice_min = mean_tau + sigma_tau * bounds[0]
    slice_max = mean_tau + sigma_tau * bounds[1]
    indx = np.where((tau >= slice_min) & (tau < slice_max))
    xs, ys = r[0][indx], r[2][indx]
    if step_profile:
        sigma = min(np.max(xs) - np.min(xs), np.max(ys) - np.min(ys)) / 2
    else:
        sigma = (np.std(xs) + np.std(ys)) / 2.
    q = np.sum(q_array)
    gamma = E_GeV / M_E_GEV
    v = np.sqrt(1 - 1 / gamma ** 2) * C_LIGHT
    B = s_to_cur(tau, sigma_tau * smooth_param, q, v)
    bunch = B[:, 1] / (q * C_LIGHT)
    x = B[:, 0]
    W = -wake_lsc(x, bunch, gamma, sigma, dz, step_profile, K_max, fill_factor) * q
    return dict(mean_tau=mean_tau, sigma_tau=sigma_tau, sigma=sigma, q=q, x=x, current=B[:, 1], W=W)


def lsc_kick(r, q_array, E_GeV, dz, **kw):
    """LSC.apply (sc.py:546-599) on a (6, N) array, in place; returns the stage dict."""
    if dz < 1e-10:
        return None
    st = lsc_stages(r, q_array, E_GeV, dz, **kw)
    dE = np.interp(r[4], st["x"], st["W"])        # the argsort of sc.py:594-596 only permutes the sum
    pc_ref = np.sqrt(E_GeV ** 2 / M_E_GEV ** 2 - 1) * M_E_GEV
    r[5] += dE * 1e-9 / pc_ref
    return st
