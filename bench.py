#!/usr/bin/env python
"""Benchmark of the space-charge kick: SC particle-kicks/s (fp64).

    python bench.py --gpus N --steps K --warmup W [--workload c2] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

One "step" = one SpaceCharge.apply over the rank's resident particles (all of SURVEY.md section 8a:
transforms, frame, mesh, deposit, Poisson solve, gather, kick).  Headline workload = BASELINE.json
configs[1]: 1M particles on a 63^3 mesh per GPU, 130 MeV, 250 pC, dz = 0.1 m.  With N > 1 GPUs every
rank holds its own shard of an N-times larger bunch (weak scaling; that is `value`), and the SAME
JSON line also carries

  north_star       BASELINE configs[2..4] at their full sizes, sharded over the N ranks:
                   C3 10M / 63^3 (redundant solve), C4 100M / 127^3 (redundant and slab-decomposed
                   solve), C5 400M / 255^3 (slab-decomposed solve), with peak device memory;
  sharded_parity   the same 10M-particle bunch kicked by one GPU alone and by the N ranks
                   (redundant and slab solve): max |row difference| / rms, charge conservation;
                   the process exits non-zero when that exceeds 1e-10.

Rank 0 prints ONE JSON line (schema: task contract).
"""
from __future__ import annotations

import argparse
import contextlib
import gc
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "SC particle-kicks/sec (fp64)"
UNIT = "particle-kicks/s"

WORKLOADS = {
    # name: (particles per GPU, mesh, description)
    "c1": (200_000, 31, "BASELINE configs[0] shape: 200k-particle Gaussian bunch, nmesh 31^3"),
    "c2": (1_000_000, 63, "BASELINE configs[1]: 1M-particle Gaussian bunch, nmesh 63^3, 130 MeV, 250 pC, single GPU"),
    "c3": (1_250_000, 63, "BASELINE configs[2] per-GPU shard: 10M particles / 8 GPUs, nmesh 63^3"),
    "c4": (12_500_000, 127, "BASELINE configs[3] per-GPU shard: 100M particles / 8 GPUs, nmesh 127^3"),
    "c5": (50_000_000, 255, "BASELINE configs[4] per-GPU shard: 400M particles / 8 GPUs, nmesh 255^3"),
}
# north-star legs of a multi-GPU run: (name, total particles, mesh, slab solve?)
NORTH_STAR = (
    ("c3_10M_63_redundant", 10_000_000, 63, False),
    ("c4_100M_127_redundant", 100_000_000, 127, False),
    ("c4_100M_127_slab", 100_000_000, 127, True),
    ("c5_400M_255_slab", 400_000_000, 255, True),
)
E_GEV, DZ, CHARGE = 0.13, 0.1, 250e-12
SIGMAS = (1e-4, 2e-5, 1e-4, 2e-5, 1e-3, 1e-4)   # generate_parray defaults (generator.py:13-14)
CHIRP = 0.01
MIN_TIMED_S = 0.5                                # the timed region lasts at least this long whatever --steps says


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=int(os.environ.get("WORLD_SIZE", "1")))
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--slab", default="auto", choices=["auto", "on", "off"],
                    help="multi-GPU Poisson solve: slab-decomposed FFT (on), redundant per rank (off), by mesh size (auto)")
    ap.add_argument("--rho-reduce", default="nvls", choices=["nvls", "nccl"],
                    help="multi-GPU charge-grid reduction: the library's in-switch multimem kernel (falls back to "
                         "NCCL without a multicast mapping) or NCCL")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-north-star", action="store_true", help="multi-GPU: skip the C3/C4/C5 legs")
    ap.add_argument("--no-parity", action="store_true", help="multi-GPU: skip the sharded-vs-single parity block")
    ap.add_argument("--legs-deadline", type=float, default=420.0,
                    help="seconds after which the multi-GPU extra legs are abandoned and the line is printed without them")
    return ap.parse_args()


# ---------------------------------------------------------------------------
# CPU arm.  bench.py may execute oracle/ only here (cpu_baseline leg and --impl reference):
#   kind "reference": the UNMODIFIED reference package staged under oracle/_ref (oracle/build_ref.py),
#                     its own SpaceCharge.apply / track();
#   kind "port":      the numpy restatement oracle/sc_oracle.py (second figure, multi-threaded FFT).
# ---------------------------------------------------------------------------
@contextlib.contextmanager
def quiet_stdout():
    """The reference prints banners on import; stdout must carry exactly one JSON line."""
    sys.stdout.flush()
    saved = os.dup(1)
    try:
        os.dup2(2, 1)
        yield
    finally:
        sys.stdout.flush()
        os.dup2(saved, 1)
        os.close(saved)


def reference_package():
    from oracle import build_ref
    if not build_ref.available():
        return None
    import logging
    with quiet_stdout():
        ref = build_ref.import_reference()
    logging.disable(logging.WARNING)
    return ref


def host_bunch(n, seed):
    """Gaussian bunch with generate_parray's default sigmas (numpy)."""
    rng = np.random.default_rng(seed)
    r = rng.standard_normal((6, n)) * np.array(SIGMAS)[:, None]
    r[5] += CHIRP * r[4] / SIGMAS[4]
    return r, np.full(n, CHARGE / n)


def reference_kicks(ref, n, mesh, kicks, warm=1):
    """Seconds per kick of the reference's own SpaceCharge.apply (sc.py:208-251) on n particles."""
    p = ref.ParticleArray(n=n)
    p.rparticles[:], p.q_array[:] = host_bunch(n, 1)
    p.E = E_GEV
    sc = ref.SpaceCharge()
    sc.nmesh_xyz = [mesh] * 3
    sc.prepare(None)
    times = []
    for i in range(warm + kicks):
        t0 = time.perf_counter()
        with quiet_stdout():
            sc.apply(p, DZ)
        if i >= warm:
            times.append(time.perf_counter() - t0)
    return float(np.median(times))


def port_kicks(n, mesh, kicks, fft, workers, warm=1):
    """Seconds per kick of the oracle port on n particles."""
    from oracle import sc_oracle as orc
    r, q = host_bunch(n, 1)
    times = []
    for i in range(warm + kicks):
        t0 = time.perf_counter()
        orc.sc_kick(r, q, E_GEV, DZ, (mesh,) * 3, fft=fft, workers=workers)
        if i >= warm:
            times.append(time.perf_counter() - t0)
    return float(np.median(times))


def fodo_lattice(ref, ncell, k1=5.0):
    seq = [ref.Marker(eid="START")]
    for i in range(ncell):
        seq += [ref.Quadrupole(l=0.2, k1=+k1, eid=f"QF{i}"), ref.Drift(l=0.3, eid=f"DA{i}"),
                ref.Quadrupole(l=0.2, k1=-k1, eid=f"QD{i}"), ref.Drift(l=0.3, eid=f"DB{i}")]
    seq.append(ref.Marker(eid="END"))
    return ref.MagneticLattice(seq)


def reference_track(ref, n, mesh, ncell):
    """The reference's own track() over `ncell` FODO cells of BASELINE config 1 (kick every 0.1 m, maps and
    get_envelope every step): (seconds, kicks)."""
    p = ref.ParticleArray(n=n)
    p.rparticles[:], p.q_array[:] = host_bunch(n, 1)
    p.E = E_GEV
    lat = fodo_lattice(ref, ncell)
    navi = ref.Navigator(lat)
    navi.unit_step = 0.1
    sc = ref.SpaceCharge()
    sc.step = 1
    sc.nmesh_xyz = [mesh] * 3
    navi.add_physics_proc(sc, lat.sequence[0], lat.sequence[-1])
    t0 = time.perf_counter()
    with quiet_stdout():
        tws, _ = ref.track(lat, p, navi, print_progress=False)
    return time.perf_counter() - t0, len(tws) - 1


def cpu_sample_size(n):
    """Bounded sample: the CPU path is linear in N at fixed mesh; cap the sample so a leg stays within ~10-30 s."""
    return min(n, 1_000_000)


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path on the host cores.  Rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    n, mesh, desc = WORKLOADS[args.workload]
    ns = cpu_sample_size(n)
    cores = os.cpu_count() or 1
    steps = max(1, min(args.steps, 6))
    ref = reference_package()
    port_sec = port_kicks(ns, mesh, min(steps, 3), "padded", cores)
    port = {"value": ns / port_sec, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{min(steps, 3)} kicks of {ns} particles on {mesh}^3, numpy restatement (oracle/sc_oracle.py) with "
                      f"scipy.fft rfftn on the padded box, workers={cores}"}
    extra = {}
    if ref is not None:
        sec = reference_kicks(ref, ns, mesh, steps, warm=1 if args.warmup else 0)
        kind, used = "reference", 1
        sample = (f"median of {steps} kicks of {ns} particles on {mesh}^3 through the unmodified reference's "
                  f"SpaceCharge.apply (oracle/_ref, ocelot {ref.__version__}); its stock path is numpy.fft + numpy ufuncs: "
                  f"single-threaded without pyFFTW / numexpr (absent from the image), host has {cores} cores; "
                  f"throughput is per particle, so the sample size does not bias it")
        tsec, tk = reference_track(ref, 200_000, 31, 2)
        extra["e2e_resident"] = {"value": 200_000 * tk / tsec, "unit": UNIT, "particles": 200_000, "kicks": tk,
                                 "seconds": tsec,
                                 "call": "reference track(lattice, p_array, navi): first 2 of the 10 FODO cells of BASELINE "
                                         "config 1 (31^3, kick every 0.1 m, maps + get_envelope every step), host numpy"}
    else:
        sec, kind, used, sample = port_sec, "port", cores, port["sample"] + " (oracle/_ref not staged)"
    value = ns / sec
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": 1 if args.warmup else 0, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": desc, "particles_per_step": ns, "nmesh": [mesh] * 3, "E_GeV": E_GEV, "dz_m": DZ},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": used, "kind": kind, "sample": sample},
        "cpu_port": port,
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    line.update(extra)
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons sampled through NVML while the GPU is under load."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index):
        self.index, self.sm, self.mask, self.smax = index, [], 0, None
        self._stop, self._t = threading.Event(), None

    def _run(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            phys = self.index
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            if vis:
                try:
                    phys = int(vis.split(",")[self.index])
                except Exception:  # noqa: BLE001
                    pass
            h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.smax = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            while not self._stop.is_set():
                self.sm.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
                try:
                    self.mask |= int(pynvml.nvmlDeviceGetCurrentClocksEventReasons(h))
                except Exception:  # noqa: BLE001
                    self.mask |= int(pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                self._stop.wait(0.002)
        except Exception:  # noqa: BLE001
            pass

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        time.sleep(0.02)
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=5)

    def summary(self):
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.smax,
                "reasons": sorted(v for k, v in self.REASONS.items() if self.mask & k), "samples": len(self.sm)}


def device_bunch(torch, n, seed, device, lo=0, hi=None, charge_total=None):
    """Synthetic Gaussian bunch with generate_parray's default sigmas, generated on the device.  With lo/hi only
    that slice of the n-particle bunch is kept (every rank draws the same stream: shards of ONE bunch)."""
    g = torch.Generator(device=device).manual_seed(seed)
    from ocelot_b200 import DeviceParticleArray
    hi = n if hi is None else hi
    p = DeviceParticleArray(hi - lo, device=device)
    r = p.rparticles
    for k in range(6):
        row = torch.randn(n, generator=g, device=device, dtype=torch.float64)
        r[k] = row[lo:hi] * SIGMAS[k]
        del row
    r[5] += CHIRP * r[4] / SIGMAS[4]
    p.q_array.fill_((CHARGE if charge_total is None else charge_total) / n)
    p.E = E_GEV
    return p


def grid_bytes(n, m, slab_world=1):
    """algorithmic grid bytes per kick, 56 M^3 + 24 n^3 (the M^3 part is split over the ranks of a
    slab-decomposed solve)."""
    return 56 * m ** 3 / slab_world + 24 * n ** 3


def bind_to_gpu_numa(index):
    """Run this process on the CPUs NVML reports as local to GPU `index`, so that the pinned host buffers
    of the end-to-end leg are first-touched on the NUMA node the GPU's PCIe root hangs off (a remote node
    costs up to 30 % of the host<->device rate).  Returns the number of CPUs bound, 0 if unavailable."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = [i * 64 + b for i, w in enumerate(mask) for b in range(64) if (int(w) >> b) & 1]
        allowed = os.sched_getaffinity(0)
        cpus = [c for c in cpus if c in allowed]
        if cpus:
            os.sched_setaffinity(0, cpus)
        return len(cpus)
    except Exception:  # noqa: BLE001
        return 0


class Ctx:
    """torch, distributed state and the shared timing helpers of one bench process."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.args = torch, dist, args
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise RuntimeError("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
        torch.cuda.set_device(self.local)
        self.device = torch.device("cuda", self.local)
        self.numa_cpus = bind_to_gpu_numa(self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.device)
        self.flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device=self.device)   # > 126 MB L2

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x):
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.device)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def timed(self, kick, steps, min_seconds=0.0, max_blocks=64):
        """Blocks of exactly `steps` kicks, per-kick CUDA events, L2 flushed (untimed) before every kick, repeated
        until the device time reaches min_seconds: (ms per kick [max over ranks], kicks timed, wall seconds)."""
        torch = self.torch
        total_ms, kicks, blocks = 0.0, 0, 0
        self.barrier()
        wall0 = time.perf_counter()
        while True:
            ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
            for a, b in ev:
                self.flush.fill_(1.0)
                a.record()
                kick()
                b.record()
            self.barrier()
            ms = self.max_over_ranks(sum(a.elapsed_time(b) for a, b in ev))     # identical on every rank
            total_ms += ms
            kicks += steps
            blocks += 1
            if total_ms * 1e-3 >= min_seconds or blocks >= max_blocks:
                break
        return total_ms / kicks, kicks, time.perf_counter() - wall0, blocks


def make_sharded(ctx, mesh, slab):
    from ocelot_b200.distributed import ShardedSpaceCharge
    s = ShardedSpaceCharge(step=1, nmesh_xyz=[mesh] * 3, slab=slab)
    s.nvls_rho = ctx.args.rho_reduce == "nvls"
    s.prepare(None)
    return s


def drop_sharded(ctx, s):
    s.finalize()
    s._engine = None
    gc.collect()
    ctx.torch.cuda.synchronize()
    ctx.torch.cuda.empty_cache()


def used_gib(torch):
    free, total = torch.cuda.mem_get_info()
    return (total - free) / 2 ** 30


def north_star_leg(ctx, name, n_total, mesh, slab, steps):
    """One BASELINE multi-GPU configuration at its full size, sharded over the ranks (strong scaling in N)."""
    from ocelot_b200 import native
    from ocelot_b200.distributed import shard_bounds
    torch = ctx.torch
    lo, hi = shard_bounds(n_total, ctx.world, ctx.rank)
    p = device_bunch(torch, hi - lo, 4321 + ctx.rank, ctx.device, charge_total=CHARGE * (hi - lo) / n_total)
    s = make_sharded(ctx, mesh, slab)
    for _ in range(3):
        s.apply(p, DZ)
    ms, kicks, _, _ = ctx.timed(lambda: s.apply(p, DZ), steps, min_seconds=0.1, max_blocks=4)
    mem = ctx.max_over_ranks(used_gib(torch))
    m = native.fft_size(mesh)
    eng = s._engine
    alg = 104 * (hi - lo) + grid_bytes(mesh, m, ctx.world if eng.slab else 1)
    peak = load_peak()[0]
    out = {"particles_total": n_total, "particles_per_gpu": hi - lo, "nmesh": [mesh] * 3, "fft_box": [m] * 3,
           "poisson": ("slab-decomposed FFT (in-switch reduce-scatter, 2 transposes fused into the FFT passes as NVLink peer "
                       "stores, phi " + ("broadcast through the switch by the inverse z pass" if getattr(eng, "mc_phi", None)
                                         is not None else "all-gather") + ")" if getattr(eng, "peer_xchg", None) is not None else
                       "slab-decomposed FFT (reduce-scatter, 2 all-to-all, all-gather)") if eng.slab
                      else "redundant per rank after the in-switch all-reduce of rho",
           "rho_reduce": "in-switch multimem kernel (NVLS)" if eng.nvls is not None else "NCCL",
           "ms_per_step": ms, "value": n_total / (ms * 1e-3), "unit": UNIT, "kicks_timed": kicks,
           "device_memory_gib_max_over_ranks": round(mem, 2),
           "whole_kick_roofline": {"algorithmic_bytes_per_gpu": alg, "achieved_gbs": alg / (ms * 1e-3) / 1e9,
                                   "frac": alg / (ms * 1e-3) / 1e9 / peak}}
    # failure detection of the exchanges that run inside the kernels: 0 = none of them ever timed out on any rank
    out["exchange_status"] = int(ctx.max_over_ranks(float(s.exchange_status())))
    drop_sharded(ctx, s)
    del p
    gc.collect()
    torch.cuda.empty_cache()
    return out


def sharded_parity(ctx, n_total=10_000_000, mesh=63):
    """The same bunch kicked by one GPU alone and by the N ranks (redundant and slab solve)."""
    from ocelot_b200 import native, DeviceParticleArray
    from ocelot_b200.distributed import shard_bounds
    torch, dist = ctx.torch, ctx.dist
    full = device_bunch(torch, n_total, 777, ctx.device)            # every rank draws the same stream
    before = full.rparticles.clone()
    solo = native.Solver(ctx.local, (mesh,) * 3)
    solo.kick_device(full.rparticles, full.q_array, E_GEV, DZ)      # "rank 0 alone" (every rank computes it for its slice)
    torch.cuda.synchronize()
    rms = full.rparticles.std(dim=1)
    lo, hi = shard_bounds(n_total, ctx.world, ctx.rank)
    out = {"particles": n_total, "nmesh": [mesh] * 3, "tolerance": 1e-10,
           "metric": "max over rows of max|sharded - single GPU| / rms(row), max over ranks"}
    worst = 0.0
    for label, slab in (("redundant", False), ("slab", True)):
        shard = DeviceParticleArray(hi - lo, device=ctx.device)
        shard.rparticles.copy_(before[:, lo:hi])
        shard.q_array.copy_(full.q_array[lo:hi])
        shard.E = E_GEV
        s = make_sharded(ctx, mesh, slab)
        s.use_graph = False
        s.apply(shard, DZ)
        torch.cuda.synchronize()
        err = float(((shard.rparticles - full.rparticles[:, lo:hi]).abs().amax(dim=1) / rms).max().item())
        moved = float(((shard.rparticles - before[:, lo:hi]).abs().amax(dim=1) / rms).max().item())
        eng = s._engine
        if eng.slab is None:
            rho_sum = float(eng.buffers["rho"].sum().item())
        else:
            t = eng.buffers["rho_slab"].sum().reshape(1).clone()
            dist.all_reduce(t)
            rho_sum = float(t.item())
        err = ctx.max_over_ranks(err)
        out[label] = {"row_error": err, "kick_size": ctx.max_over_ranks(moved),
                      "charge_conservation": abs(rho_sum / CHARGE - 1.0),
                      "solve": "slab-decomposed" if eng.slab else "redundant"}
        status = int(ctx.max_over_ranks(float(s.exchange_status())))
        out[label]["exchange_status"] = status
        worst = max(worst, err, abs(rho_sum / CHARGE - 1.0), float(status))
        drop_sharded(ctx, s)
        del shard
    out["ok"] = bool(worst < 1e-10)
    del full, before, solo
    gc.collect()
    torch.cuda.empty_cache()
    return out


def load_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
        return float(peaks["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (driver-measured copy bandwidth)"
    except Exception:  # noqa: BLE001
        return 6650.0, "fallback 6650 GB/s (B200_PROFILING.md)"


def resident_track_leg(ctx, n, mesh=31):
    """BASELINE config 1 (10 m FODO, 100 kicks, maps + beam moments every step) through the package's resident
    tracking loop: one H2D of the bunch at the start, one D2H at the end, both inside the timed region; the maps are
    the ones the reference's Navigator produced for this lattice (tests/golden/track_c1.npz, recorded by
    oracle/make_golden.py), applied by the device map kernel."""
    torch = ctx.torch
    from ocelot_b200 import SpaceCharge, DeviceParticleArray, EnvelopeRecorder
    from ocelot_b200.track import replay_track
    path = os.path.join(ROOT, "tests", "golden", "track_c1.npz")
    with np.load(path) as z:
        R, B, map_step, kick_dz = z["R"], z["B"], z["map_step"], z["kick_dz"]
    r_host, q_host = host_bunch(n, 5)
    hr = torch.from_numpy(r_host).pin_memory()
    hq = torch.from_numpy(q_host).pin_memory()
    out_r = torch.empty_like(hr).pin_memory()
    sc = SpaceCharge(step=1, nmesh_xyz=[mesh] * 3, device=ctx.local)
    sc.prepare(None)
    moments = []

    def once():
        p = DeviceParticleArray(n, device=ctx.device)
        p.rparticles.copy_(hr, non_blocking=True)
        p.q_array.copy_(hq, non_blocking=True)
        p.E = E_GEV
        # as ocelot_b200.track.track does: the moments of every step are recorded on the device (two moment passes per
        # step, no host synchronisation) and read back once, after the loop
        rec = EnvelopeRecorder(ctx.device)
        rec.record(p)
        replay_track(p, R, B, map_step, kick_dz, sc, after_step=lambda step, pa: rec.record(pa))
        out_r.copy_(p.rparticles, non_blocking=True)
        moments[:] = rec.collect()
        torch.cuda.synchronize()
        return p

    once()
    best = float("inf")
    for _ in range(3):
        t0 = time.perf_counter()
        once()
        best = min(best, time.perf_counter() - t0)
    kicks = int(np.count_nonzero(kick_dz))
    return {"particles": n, "nmesh": [mesh] * 3, "kicks": kicks, "maps": int(len(R)), "seconds": best,
            "value": n * kicks / best, "unit": UNIT, "ms_per_step": best / kicks * 1e3,
            "h2d_bytes": 56 * n, "d2h_bytes": 48 * n + 24 * 8 * 128 * ((kicks + 1 + 127) // 128),
            "final_sigma_x": float(np.sqrt(moments[-1].xx))}


def run_native(args):
    ctx = Ctx(args)
    torch, dist = ctx.torch, ctx.dist
    from ocelot_b200 import native
    world, rank, local, device = ctx.world, ctx.rank, ctx.local, ctx.device
    n, mesh, desc = WORKLOADS[args.workload]
    p = device_bunch(torch, n, 1234 + rank, device)
    r, q = p.rparticles, p.q_array
    sharded = None
    if world > 1:
        # staged kick + collectives, captured into one CUDA graph (ocelot_b200/distributed.py)
        sharded = make_sharded(ctx, mesh, {"auto": None, "on": True, "off": False}[args.slab])
        sharded.use_graph = False
        sharded.apply(p, DZ)
        solver = sharded._engine.solver
        l0 = solver.launch_count()
        sharded.apply(p, DZ)
        launches_per_kick = solver.launch_count() - l0
        sharded.use_graph = True
    else:
        solver = native.Solver(local, (mesh,) * 3)     # whole kick = one CUDA graph inside the library
        launches_per_kick = None

    def kick():
        if sharded is not None:
            sharded.apply(p, DZ)
        else:
            solver.kick_device(r, q, E_GEV, DZ)

    for _ in range(max(3, args.warmup)):
        kick()
    ctx.barrier()

    # ---- timed region: blocks of K steps, per-step events, L2 flushed (untimed) between steps, >= 0.5 s ----
    K = args.steps
    launches0 = solver.launch_count()
    with ClockSampler(local) as clocks:
        ms_per_step, kicks_timed, wall, blocks = ctx.timed(kick, K, MIN_TIMED_S)
    launches = solver.launch_count() - launches0
    if launches_per_kick is not None:
        launches = launches_per_kick * kicks_timed
    value = world * n / (ms_per_step * 1e-3)

    # ---- warm-L2 figure (no flush), reported beside the headline ----
    ctx.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(K):
        kick()
    b.record()
    ctx.barrier()
    warm_ms = ctx.max_over_ranks(a.elapsed_time(b) / K)

    # ---- dominant kernel live timing: the library's own events around each stage ----
    if sharded is not None:
        sharded.use_graph = False
    solver.enable_timers(True)
    acc = {}
    kt = max(K, 20)
    for _ in range(kt):
        ctx.flush.fill_(1.0)
        kick()
        for k_, v_ in solver.timers().items():
            acc[k_] = acc.get(k_, 0.0) + v_ / kt
    solver.enable_timers(False)
    if sharded is not None:
        sharded.use_graph = True
    ctx.barrier()

    line = None
    if rank == 0:
        peak, peak_src = load_peak()
        m = native.fft_size(mesh)
        slab_on = sharded is not None and sharded._engine.slab is not None
        roof = None
        if acc:
            dom = "kick"      # k_gather_kick: trilinear gather + kick + back-transform, 6 rows in, 6 rows out
            bytes_dom = 96 * n
            t_dom = acc[dom] * 1e-3
            traffic = None
            try:
                with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                    traffic = json.load(f).get(args.workload, {}).get("k_gather_kick")
            except Exception:  # noqa: BLE001
                pass
            ach = bytes_dom / t_dom / 1e9
            alg = 104 * n + grid_bytes(mesh, m, world if slab_on else 1)
            roof = {"bound": "hbm", "kernel": "k_gather_kick", "achieved": ach, "peak": peak, "unit": "GB/s",
                    "frac": ach / peak, "traffic": traffic, "algorithmic_bytes_per_launch": bytes_dom,
                    "kernel_ms": acc[dom], "peak_source": peak_src,
                    "stage_ms": {k_: round(v_, 4) for k_, v_ in acc.items()},
                    "whole_kick": {"algorithmic_bytes": alg, "achieved": alg / (ms_per_step * 1e-3) / 1e9,
                                   "frac": alg / (ms_per_step * 1e-3) / 1e9 / peak}}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": max(3, args.warmup),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": desc, "particles_per_gpu": n, "particles_total": n * world, "nmesh": [mesh] * 3,
                       "fft_box": [m] * 3, "E_GeV": E_GEV, "dz_m": DZ, "parallelism": f"particle-shard x{world}",
                       "collectives_per_kick": 0 if world == 1 else (6 if slab_on else 3),
                       "poisson": "single GPU" if world == 1 else ("slab-decomposed FFT (reduce-scatter, 2 all-to-all, all-gather)"
                                                                   if slab_on else "redundant per rank after all-reduce of rho"),
                       "rho_reduce": "n/a" if world == 1 else ("in-switch multimem kernel (NVLS)" if sharded._engine.nvls
                                                               is not None else "NCCL"),
                       "scalar_exchanges": "n/a" if world == 1 else ("inside the sweep kernels over NVLink peer memory"
                                                                     if sharded._engine.mailbox is not None else "NCCL"),
                       "cuda_graph": "whole kick captured once; parameter node refreshed per kick",
                       "l2": "256 MiB buffer written between timed steps (untimed); per-step CUDA events",
                       "timed": f"{blocks} block(s) of {K} steps = {kicks_timed} kicks, at least {MIN_TIMED_S} s of device time"},
            "timed_kicks": kicks_timed,
            "warm_l2": {"ms_per_step": warm_ms, "value": world * n / (warm_ms * 1e-3)},
            "gpu_launches": int(launches), "wall_s_timed_region": wall,
            "clocks": clocks.summary(), "roofline": roof,
        }

    # ---- end-to-end through the plugin call with HOST (pinned) buffers ----
    if not args.no_e2e:
        from ocelot_b200 import SpaceCharge, ParticleArray
        sc = SpaceCharge(step=1, nmesh_xyz=[mesh] * 3, device=local)
        sc.prepare(None)
        hp = ParticleArray(0)
        host_r = torch.empty((6, n), dtype=torch.float64).pin_memory()
        host_q = torch.empty(n, dtype=torch.float64).pin_memory()
        host_r.copy_(r.cpu())
        host_q.copy_(q.cpu())
        hp.rparticles, hp.q_array, hp.E = host_r.numpy(), host_q.numpy(), E_GEV
        ke = max(3, min(K, 10))
        for _ in range(2):
            sc.apply(hp, DZ)
        ctx.barrier()
        dt = float("inf")
        for _ in range(3):             # best of 3 blocks of ke kicks: host-side jitter (page faults, NUMA) is one-sided
            t0 = time.perf_counter()
            for _ in range(ke):
                sc.apply(hp, DZ)       # H2D of 6 rows + q, kick, D2H of 6 rows, synchronous
            torch.cuda.synchronize()
            dt = min(dt, (time.perf_counter() - t0) / ke)
        dt = ctx.max_over_ranks(dt)
        if line is not None:
            line["e2e"] = {"value": world * n / dt, "unit": UNIT, "h2d_bytes_per_step": 56 * n,
                           "d2h_bytes_per_step": 48 * n, "ms_per_step": dt * 1e3,
                           "call": "ocelot_b200.SpaceCharge.apply(p_array, dz) on pinned host arrays "
                                   "(independent replica per rank); best of 3 blocks of %d kicks" % ke,
                           "cpus_local_to_gpu": ctx.numa_cpus}
        del sc, hp, host_r, host_q

    if world == 1 and line is not None:
        # ---- resident tracking loop, end to end (config 1: maps + kick + moments per step) ----
        if not args.no_e2e:
            try:
                line["e2e_resident"] = {
                    "call": "ocelot_b200.track.replay_track + beam moments every step (EnvelopeRecorder, as "
                            "ocelot_b200.track.track does: moments kept on the device, read back once after the loop): "
                            "BASELINE config-1 lattice (10 m FODO, 100 kicks, 31^3), bunch H2D once at the start, bunch "
                            "and moments D2H once at the end, all inside the timed region; device transfer maps; "
                            "best of 3 runs",
                    "runs": [resident_track_leg(ctx, 200_000), resident_track_leg(ctx, 1_000_000)]}
            except Exception as exc:  # noqa: BLE001
                line["e2e_resident"] = {"error": repr(exc)}
        # ---- the 1-D sibling on the same resident bunch, reported beside the headline ----
        from ocelot_b200 import LSC
        import types
        dp = types.SimpleNamespace(rparticles=r, q_array=q, E=E_GEV)     # apply() duck-types these three
        lsc = LSC(step=1, device=local)
        for _ in range(3):
            lsc.apply(dp, DZ)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        kl = max(3, min(K, 20))
        a.record()
        for _ in range(kl):
            lsc.apply(dp, DZ)
        b.record()
        torch.cuda.synchronize()
        lms = a.elapsed_time(b) / kl
        line["lsc"] = {"value": n / (lms * 1e-3), "unit": "LSC particle-kicks/s", "ms_per_step": lms,
                       "grid_points": int(lsc.last_params["nb"]),
                       "note": "ocelot_b200.LSC.apply on the resident bunch (grid derived on the device, no host "
                               "synchronisation); not part of `value`"}

    # ---- CPU baseline on the host cores (rank 0, N = 1 only) ----
    if line is not None and world == 1 and not args.no_cpu_baseline:
        ns = cpu_sample_size(n)
        cores = os.cpu_count() or 1
        ref = reference_package()
        if ref is not None:
            sec = reference_kicks(ref, ns, mesh, 3)
            line["cpu_baseline"] = {"value": ns / sec, "unit": UNIT, "cores": 1, "kind": "reference",
                                    "sample": f"median of 3 kicks (1 warm-up) of {ns} particles on {mesh}^3 through the "
                                              f"unmodified reference's SpaceCharge.apply (oracle/_ref, ocelot "
                                              f"{ref.__version__}; numpy.fft on the (2n-1)^3 box, single-threaded "
                                              f"without pyFFTW / numexpr; host has {cores} cores)"}
        else:
            sec = port_kicks(ns, mesh, 3, "reference", 1)
            line["cpu_baseline"] = {"value": ns / sec, "unit": UNIT, "cores": 1, "kind": "port",
                                    "sample": f"median of 3 kicks (1 warm-up) of {ns} particles on {mesh}^3, oracle port "
                                              f"in the reference's own configuration (oracle/_ref not staged)"}

    # ---- multi-GPU: sharded-vs-single parity and the north-star configurations, in the same line ----
    rc = 0
    if world > 1:
        drop_sharded(ctx, sharded)
        del p, r, q
        gc.collect()
        torch.cuda.empty_cache()
        state = {"line": line, "printed": False}

        def emit():
            if state["printed"]:
                return
            state["printed"] = True
            if state["line"] is not None:
                print(json.dumps(state["line"]), flush=True)

        def deadline():
            # a leg hung (or is far slower than planned): keep the headline, say so, leave
            if line is not None:
                line.setdefault("north_star", {})["abandoned"] = f"legs exceeded {args.legs_deadline:.0f} s"
            emit()
            os._exit(0)

        timer = threading.Timer(args.legs_deadline, deadline)
        timer.daemon = True
        timer.start()
        if not args.no_parity:
            par = sharded_parity(ctx)
            if line is not None:
                line["sharded_parity"] = par
            if not par["ok"]:
                rc = 3
        if not args.no_north_star:
            legs = {}
            if line is not None:
                line["north_star"] = legs
            for name, n_total, nm, slab in NORTH_STAR:
                t0 = time.perf_counter()
                leg = north_star_leg(ctx, name, n_total, nm, slab, max(3, min(K, 10)))
                leg["leg_wall_s"] = round(time.perf_counter() - t0, 1)
                legs[name] = leg
        timer.cancel()
        emit()
        torch.cuda.synchronize()
        dist.destroy_process_group()
    elif line is not None:
        print(json.dumps(line), flush=True)
    if rc:
        sys.exit(rc)


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_native(a)
