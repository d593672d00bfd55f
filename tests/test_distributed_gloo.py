"""world_size-2 gloo test of the particle-sharded kick (CPU only).

The sharding / collective logic of ocelot_b200.distributed is exercised with a
stage engine backed by the oracle (test infrastructure standing in for the
CUDA stages): two ranks, each holding half of the bunch, must reproduce the
single-process oracle kick."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import sc_oracle as orc
from ocelot_b200.distributed import shard_bounds, sharded_kick


class OracleStageEngine:
    """The five stages restated on the CPU from oracle functions, with the same
    reduction buffers as the native handle (include/ocelot_sc.h)."""

    def __init__(self, nmesh):
        self.nmesh = np.array(nmesh)
        self.buffers = {"momentum": torch.zeros(4, dtype=torch.float64),
                        "extent_max": torch.zeros(6, dtype=torch.float64),
                        "extent_sum": torch.zeros(4, dtype=torch.float64),
                        "rho": torch.zeros(int(np.prod(nmesh)), dtype=torch.float64)}

    def momentum(self, r, q, E):
        self.gamref = E / orc.M_E_GEV
        self.xp = orc.mad_to_cartesian(r.numpy(), self.gamref)
        self.buffers["momentum"][:3] = torch.from_numpy(self.xp[3:6].sum(axis=1))
        self.buffers["momentum"][3] = r.shape[1]

    def _frame(self):
        s = self.buffers["momentum"].numpy()
        mean = s[:3] / s[3]
        # bunch_frame takes the momenta; feed it the global mean as a single "particle"
        return orc.bunch_frame(mean.reshape(3, 1))

    def extent(self, r, q, E):
        self.T, self.pav, self.gamma0, self.beta0 = self._frame()
        X = np.dot(self.xp[0:3].T, self.T)
        X[:, 2] *= self.gamma0
        self.X = X
        qn = q.numpy()
        self.buffers["extent_max"][:3] = torch.from_numpy(X.max(axis=0))
        self.buffers["extent_max"][3:] = torch.from_numpy(-X.min(axis=0))
        self.buffers["extent_sum"][:3] = torch.from_numpy(qn @ X)
        self.buffers["extent_sum"][3] = qn.sum()

    def _mesh(self, draws):
        em, es = self.buffers["extent_max"].numpy(), self.buffers["extent_sum"].numpy()
        extent = em[:3] + em[3:]
        if draws is not None:
            extent = extent * draws[0]
        steps = extent / (self.nmesh - 3)
        xmid = (es[:3] / steps) / es[3]
        xoff = np.floor(-em[3:] / steps - xmid) + xmid
        if draws is not None:
            xoff = xoff + draws[1]
        return steps, xoff

    def deposit(self, r, q, E, draws):
        self.steps, self.xoff = self._mesh(draws)
        self.Xg = self.X / self.steps - self.xoff
        idx = orc.cell_index(self.Xg, self.nmesh)
        self.buffers["rho"][:] = torch.from_numpy(orc.deposit_ngp(idx, q.numpy(), self.nmesh).ravel())

    def solve(self, draws):
        rho = self.buffers["rho"].numpy().reshape(tuple(self.nmesh))
        phi = orc.poisson_potential(rho, self.steps, fft="padded")
        self.E3 = orc.staggered_field(phi, self.steps)

    def kick(self, r, E, dz, draws):
        Xg, g0 = self.Xg, self.gamma0
        Ex = orc.trilinear(self.E3[0], Xg[:, 0], Xg[:, 1] + 0.5, Xg[:, 2] + 0.5) * g0
        Ey = orc.trilinear(self.E3[1], Xg[:, 0] + 0.5, Xg[:, 1], Xg[:, 2] + 0.5) * g0
        Ez = orc.trilinear(self.E3[2], Xg[:, 0] + 0.5, Xg[:, 1] + 0.5, Xg[:, 2])
        xp = self.xp
        betref = np.sqrt(1 - self.gamref ** -2)
        cdT = dz / betref
        p = np.dot(xp[3:6].T, self.T).T
        p[0] += cdT * (1 - self.beta0 * self.beta0) * Ex
        p[1] += cdT * (1 - self.beta0 * self.beta0) * Ey
        p[2] += cdT * Ez
        xp[3:6] = np.dot(p.T, self.T.T).T
        orc.cartesian_to_mad(xp, r.numpy(), self.gamref)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, nmesh, draws, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        np.random.seed(3)
        r0, q0, E = orc.gaussian_bunch(n, energy=0.05, charge=1e-10)
        q0 = q0 * (0.5 + np.random.rand(n))
        lo, hi = shard_bounds(n, world, rank)
        r = torch.from_numpy(r0[:, lo:hi].copy())
        q = torch.from_numpy(q0[lo:hi].copy())
        sharded_kick(OracleStageEngine(nmesh), r, q, E, 0.07, draws)
        out[rank] = (lo, hi, r.numpy().copy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("draws", [None, (1.05, 0.3)])
def test_two_rank_sharded_kick_matches_single_process(draws):
    n, nmesh, world = 6001, (15, 13, 17), 2
    port = _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, port, n, nmesh, draws, out), nprocs=world, join=True)
        parts = [out[k] for k in range(world)]
    np.random.seed(3)
    r0, q0, E = orc.gaussian_bunch(n, energy=0.05, charge=1e-10)
    q0 = q0 * (0.5 + np.random.rand(n))
    ref = r0.copy()
    kw = {} if draws is None else dict(mesh_scale=draws[0], mesh_shift=draws[1])
    orc.sc_kick(ref, q0, E, 0.07, nmesh, fft="padded", **kw)
    got = np.empty_like(ref)
    covered = 0
    for lo, hi, rr in parts:
        got[:, lo:hi] = rr
        covered += hi - lo
    assert covered == n
    for row in range(6):
        assert np.max(np.abs(got[row] - ref[row])) / np.std(ref[row]) < 1e-10


class DeferredOracleEngine(OracleStageEngine):
    """The same double with the hooks of the native engine's NCCL-fallback form (begin_kick, finish_momentum,
    finish_extent: geometry derived only AFTER the caller's all-reduce), recording the order of the calls."""

    def __init__(self, nmesh):
        super().__init__(nmesh)
        self.log = []

    def begin_kick(self, draws, multi):
        self.log.append(("begin", multi))

    def momentum(self, r, q, E):
        self.log.append("momentum")
        super().momentum(r, q, E)
        self.local_count = float(self.buffers["momentum"][3])

    def finish_momentum(self):
        # must run after the all-reduce: the count is now the whole bunch's
        assert float(self.buffers["momentum"][3]) > self.local_count
        self.log.append("finish_momentum")

    def extent(self, r, q, E):
        self.log.append("extent")
        super().extent(r, q, E)
        self.local_sumq = float(self.buffers["extent_sum"][3])

    def finish_extent(self):
        assert float(self.buffers["extent_sum"][3]) > self.local_sumq
        self.log.append("finish_extent")

    def deposit(self, r, q, E, draws):
        self.log.append("deposit")
        super().deposit(r, q, E, draws)


def _deferred_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        np.random.seed(3)
        r0, q0, E = orc.gaussian_bunch(4000, energy=0.05, charge=1e-10)
        lo, hi = shard_bounds(4000, world, rank)
        eng = DeferredOracleEngine((9, 9, 9))
        sharded_kick(eng, torch.from_numpy(r0[:, lo:hi].copy()), torch.from_numpy(q0[lo:hi].copy()), E, 0.05, None)
        out[rank] = list(eng.log)
    finally:
        dist.destroy_process_group()


def test_deferred_finish_hooks_run_after_the_collectives():
    """Without a peer-memory mailbox the sweeps only reduce locally; the engine's finish hooks (ocl_sc_stage_finish in
    the native engine) must be called after each all-reduce and before the next stage."""
    port = _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_deferred_worker, args=(2, port, out), nprocs=2, join=True)
        logs = [out[k] for k in range(2)]
    for log in logs:
        assert log == [("begin", True), "momentum", "finish_momentum", "extent", "finish_extent", "deposit"]


def test_shard_bounds_cover_everything():
    for n in (0, 1, 7, 1000, 1001):
        for w in (1, 2, 3, 8):
            spans = [shard_bounds(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


# ---------------------------------------------------------------------------
# particle-sharded longitudinal space charge
# ---------------------------------------------------------------------------
class OracleLscEngine:
    """The device stages of the LSC kick restated with numpy (test double for NativeLscEngine):
    integer fixed-point histogram like csrc/sc_lsc.cu, so the all-reduce is exact."""
    SHIFT = 40

    def stats(self, r, q):
        r, q = r.numpy(), q.numpy()
        tau = r[4]
        m = tau.mean()
        return torch.tensor([tau.size, m, np.sum((tau - m) ** 2), tau.min(), tau.max(), q.sum(), r[0].sum(),
                             r[2].sum()], dtype=torch.float64)

    def deposit(self, r, params):
        from oracle import lsc_oracle as lo
        r = r.numpy()
        nb = int(params["nb"])
        C = lo.cic_counts(r[4], params["a"], params["ds"], nb)
        self.bins = torch.from_numpy(np.rint(C * 2.0 ** self.SHIFT).astype(np.int64))
        sl = (r[4] >= params["slice_min"]) & (r[4] < params["slice_max"])
        x, y = r[0][sl], r[2][sl]
        dx, dy = x - params["x_shift"], y - params["y_shift"]
        self.smax = torch.tensor([x.max(), -x.min(), y.max(), -y.min()], dtype=torch.float64)
        self.ssum = torch.tensor([x.size, dx.sum(), (dx * dx).sum(), dy.sum(), (dy * dy).sum()], dtype=torch.float64)
        return self.bins, self.smax, self.ssum

    def solve_kick(self, r, params):
        from oracle import lsc_oracle as lo
        r = r.numpy()
        nb, a, ds, K = int(params["nb"]), params["a"], params["ds"], int(params["K"])
        C = self.bins.numpy().astype(np.float64) / 2.0 ** self.SHIFT
        if K >= 0:
            G = np.exp(-0.5 * (np.arange(-K, K + 1) * ds / params["sigma_s"]) ** 2)
            C = np.convolve(C, G / G.sum())[K:nb + K]
        s = self.ssum.numpy()
        if params["step_profile"]:
            m = self.smax.numpy()
            sigma = min(m[0] + m[1], m[2] + m[3]) / 2
        else:
            sigma = (np.sqrt(s[2] / s[0] - (s[1] / s[0]) ** 2) + np.sqrt(s[4] / s[0] - (s[3] / s[0]) ** 2)) / 2
        x = np.arange(nb) * ds + a
        bunch = params["v"] * C / (ds * C.sum()) / lo.C_LIGHT
        W = -lo.wake_lsc(x, bunch, params["gamma"], sigma, params["dz"], bool(params["step_profile"])) * params["q"]
        r[5] += np.interp(r[4], x, W) * 1e-9 / params["pc_ref"]


def _lsc_worker(rank, world, port, n, step_profile, out):
    from ocelot_b200 import LSC
    from ocelot_b200.distributed import sharded_lsc_kick
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        np.random.seed(3)
        r0, q0, E = orc.gaussian_bunch(n, energy=0.05, charge=1e-10)
        r0[0] += 2e-4
        lo_, hi = shard_bounds(n, world, rank)
        r = torch.from_numpy(r0[:, lo_:hi].copy())
        q = torch.from_numpy(q0[lo_:hi].copy())
        sharded_lsc_kick(OracleLscEngine(), LSC(step_profile=step_profile), r, q, E, 0.3)
        out[rank] = (lo_, hi, r.numpy().copy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("step_profile", [False, True])
def test_two_rank_sharded_lsc_matches_single_process(step_profile):
    from oracle import lsc_oracle as lo
    n, world = 20001, 2
    port = _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_lsc_worker, args=(world, port, n, step_profile, out), nprocs=world, join=True)
        parts = [out[k] for k in range(world)]
    np.random.seed(3)
    r0, q0, E = orc.gaussian_bunch(n, energy=0.05, charge=1e-10)
    r0[0] += 2e-4
    ref = r0.copy()
    lo.lsc_kick(ref, q0, E, 0.3, step_profile=step_profile)
    got = np.empty_like(ref)
    for lo_, hi, rr in parts:
        got[:, lo_:hi] = rr
    d_ref = ref[5] - r0[5]
    assert np.abs((got[5] - r0[5]) - d_ref).max() <= 1e-10 * np.abs(d_ref).max()
    assert np.array_equal(got[:5], r0[:5])


def test_combine_lsc_stats_is_the_global_statistic():
    from ocelot_b200.distributed import combine_lsc_stats
    rng = np.random.default_rng(0)
    tau = rng.normal(1e-3, 2e-4, 1000)
    parts = np.split(tau, [100, 350, 351])
    rows = [[p.size, p.mean(), np.sum((p - p.mean()) ** 2), p.min(), p.max(), p.size * 1.0, 0.0, 0.0] for p in parts]
    rows.append([0, 0, 0, 0, 0, 0, 0, 0])                      # an empty shard
    st = combine_lsc_stats(rows)
    assert st["n"] == 1000 and st["mean_tau"] == pytest.approx(tau.mean(), rel=1e-14)
    assert np.sqrt(st["m2_tau"] / st["n"]) == pytest.approx(np.std(tau), rel=1e-13)
    assert (st["min_tau"], st["max_tau"]) == (tau.min(), tau.max())
