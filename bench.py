#!/usr/bin/env python
"""Benchmark of the space-charge kick: SC particle-kicks/s (fp64).

    python bench.py --gpus N --steps K --warmup W [--workload c2] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

One "step" = one SpaceCharge.apply over the rank's resident particles (all of
SURVEY.md section 8a: transforms, frame, mesh, deposit, Poisson solve, gather,
kick).  Default workload = BASELINE.json configs[1]: 1M particles on a 63^3 mesh
per GPU, 130 MeV, 250 pC, dz = 0.1 m.  With N > 1 GPUs every rank holds its own
shard of an N-times larger bunch (weak scaling); the ranks all-reduce three
small buffers and the charge grid per kick and solve redundantly
(ocelot_b200/distributed.py).

Rank 0 prints ONE JSON line (schema: see the task contract in DESIGN.md section 7).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
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
E_GEV, DZ, CHARGE = 0.13, 0.1, 250e-12
SIGMAS = (1e-4, 2e-5, 1e-4, 2e-5, 1e-3, 1e-4)   # generate_parray defaults (generator.py:13-14)
CHIRP = 0.01


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
    return ap.parse_args()


# ---------------------------------------------------------------------------
# CPU arm: the oracle port of the reference algorithm (bench.py may execute oracle/
# only here: cpu_baseline leg and --impl reference)
# ---------------------------------------------------------------------------
def cpu_bunch(n, seed):
    from oracle import sc_oracle as orc
    np.random.seed(seed)
    return orc.gaussian_bunch(n, energy=E_GEV, charge=CHARGE)


def cpu_kicks(n, mesh, kicks, fft, workers, warm=1):
    """Seconds per kick of the oracle port on n particles."""
    from oracle import sc_oracle as orc
    r, q, E = cpu_bunch(n, 1)
    for _ in range(warm):
        orc.sc_kick(r, q, E, DZ, (mesh,) * 3, fft=fft, workers=workers)
    times = []
    for _ in range(kicks):
        t0 = time.perf_counter()
        orc.sc_kick(r, q, E, DZ, (mesh,) * 3, fft=fft, workers=workers)
        times.append(time.perf_counter() - t0)
    return float(np.median(times))


def cpu_sample_size(n):
    """Bounded sample: the CPU path is linear in N at fixed mesh; cap the sample so the
    leg stays within ~10-30 s on one core."""
    return min(n, 1_000_000)


def run_reference(args):
    """--impl reference: the reference's algorithm on the host cores (oracle port; the
    Python reference itself cannot travel to the GPU box).  Rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n, mesh, desc = WORKLOADS[args.workload]
    ns = cpu_sample_size(n)
    cores = os.cpu_count() or 1
    per_kick = []
    for _ in range(max(1, args.warmup and 1)):
        cpu_kicks(ns, mesh, 1, "padded", cores, warm=0)
    steps = max(1, min(args.steps, 8))
    for _ in range(steps):
        per_kick.append(cpu_kicks(ns, mesh, 1, "padded", cores, warm=0))
    sec = float(np.median(per_kick))
    value = ns / sec
    sample = (f"{steps} kicks of {ns} particles on {mesh}^3 (same bunch parameters), numpy port with "
              f"scipy.fft rfftn workers={cores}; throughput is per particle so the sample size does not bias it")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": 1, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": desc, "particles_per_step": ns, "nmesh": [mesh] * 3, "E_GeV": E_GEV, "dz_m": DZ},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
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


def device_bunch(torch, n, seed, device):
    """Synthetic Gaussian bunch with generate_parray's default sigmas, generated on the device."""
    g = torch.Generator(device=device).manual_seed(seed)
    from ocelot_b200 import DeviceParticleArray
    p = DeviceParticleArray(n, device=device)
    r = p.rparticles
    for k in range(6):
        r[k] = torch.randn(n, generator=g, device=device, dtype=torch.float64) * SIGMAS[k]
    r[5] += CHIRP * r[4] / SIGMAS[4]
    p.q_array.fill_(CHARGE / n)
    p.E = E_GEV
    return p


def grid_bytes(n, m):
    """algorithmic grid bytes per kick, 56 M^3 + 24 n^3."""
    return 56 * m ** 3 + 24 * n ** 3


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


def run_native(args):
    import torch
    import torch.distributed as dist
    from ocelot_b200 import native
    from ocelot_b200.distributed import ShardedSpaceCharge

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    numa_cpus = bind_to_gpu_numa(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    n, mesh, desc = WORKLOADS[args.workload]
    p = device_bunch(torch, n, 1234 + rank, device)
    r, q = p.rparticles, p.q_array
    sharded = None
    if world > 1:
        # staged kick + NCCL collectives, captured into one CUDA graph (ocelot_b200/distributed.py)
        sharded = ShardedSpaceCharge(step=1, nmesh_xyz=[mesh] * 3,
                                     slab={"auto": None, "on": True, "off": False}[args.slab])
        sharded.nvls_rho = args.rho_reduce == "nvls"
        sharded.prepare(None)
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

    flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device=device)   # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(3, args.warmup)):
        kick()
    barrier()

    # ---- timed region: K steps, per-step events, L2 flushed (untimed) between steps ----
    K = args.steps
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    launches0 = solver.launch_count()
    with ClockSampler(local) as clocks:
        barrier()
        wall0 = time.perf_counter()
        for a, b in ev:
            flush.fill_(1.0)
            a.record()
            kick()
            b.record()
        barrier()
        wall = time.perf_counter() - wall0
    launches = solver.launch_count() - launches0
    if launches_per_kick is not None:
        launches = launches_per_kick * K
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    t = torch.tensor([dev_ms], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    ms_per_step = total_ms / K
    value = world * n * K / (total_ms * 1e-3)

    # ---- warm-L2 figure (no flush), reported beside the headline ----
    barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(K):
        kick()
    b.record()
    barrier()
    warm_ms = a.elapsed_time(b) / K

    # ---- dominant kernel live timing: the library's own events around each stage ----
    if sharded is not None:
        sharded.use_graph = False
    solver.enable_timers(True)
    acc = {}
    for _ in range(K):
        flush.fill_(1.0)
        kick()
        for k_, v_ in solver.timers().items():
            acc[k_] = acc.get(k_, 0.0) + v_ / K
    solver.enable_timers(False)
    barrier()

    line = None
    if rank == 0:
        peaks, peak_src = {}, "fallback 6650 GB/s (B200_PROFILING.md)"
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                peaks = json.load(f)
            peak_src = "MEASURED_PEAKS.json hbm_gbs (driver-measured copy bandwidth)"
        except Exception:  # noqa: BLE001
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        m = native.fft_size(mesh)
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
            roof = {"bound": "hbm", "kernel": "k_gather_kick", "achieved": ach, "peak": peak, "unit": "GB/s",
                    "frac": ach / peak, "traffic": traffic, "algorithmic_bytes_per_launch": bytes_dom,
                    "kernel_ms": acc[dom], "peak_source": peak_src,
                    "stage_ms": {k_: round(v_, 4) for k_, v_ in acc.items()},
                    "whole_kick": {"algorithmic_bytes": 104 * n + grid_bytes(mesh, m),
                                   "achieved": (104 * n + grid_bytes(mesh, m)) / (ms_per_step * 1e-3) / 1e9,
                                   "frac": (104 * n + grid_bytes(mesh, m)) / (ms_per_step * 1e-3) / 1e9 / peak}}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": max(3, args.warmup),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": desc, "particles_per_gpu": n, "particles_total": n * world, "nmesh": [mesh] * 3,
                       "fft_box": [m] * 3, "E_GeV": E_GEV, "dz_m": DZ, "parallelism": f"particle-shard x{world}",
                       "collectives_per_kick": 0 if world == 1 else (6 if (sharded is not None and sharded._engine.slab) else 3),
                       "poisson": "single GPU" if world == 1 else ("slab-decomposed FFT (reduce-scatter, 2 all-to-all, all-gather)"
                                                                   if sharded._engine.slab else "redundant per rank after all-reduce of rho"),
                       "rho_reduce": "n/a" if world == 1 else ("in-switch multimem kernel (NVLS)" if sharded._engine.nvls
                                                               is not None else "NCCL"),
                       "cuda_graph": "whole kick captured once; parameter node refreshed per kick",
                       "l2": "256 MiB buffer written between timed steps (untimed); per-step CUDA events"},
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
        barrier()
        dt = float("inf")
        for _ in range(3):             # best of 3 blocks of ke kicks: host-side jitter (page faults, NUMA) is one-sided
            t0 = time.perf_counter()
            for _ in range(ke):
                sc.apply(hp, DZ)       # H2D of 6 rows + q, kick, D2H of 6 rows, synchronous
            torch.cuda.synchronize()
            dt = min(dt, (time.perf_counter() - t0) / ke)
        te = torch.tensor([dt], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        if line is not None:
            line["e2e"] = {"value": world * n / float(te.item()), "unit": UNIT, "h2d_bytes_per_step": 56 * n,
                           "d2h_bytes_per_step": 48 * n, "ms_per_step": float(te.item()) * 1e3,
                           "call": "ocelot_b200.SpaceCharge.apply(p_array, dz) on pinned host arrays "
                                   "(independent replica per rank); best of 3 blocks of %d kicks" % ke,
                           "cpus_local_to_gpu": numa_cpus}

    # ---- the 1-D sibling on the same resident bunch, reported beside the headline ----
    if line is not None and world == 1:
        from ocelot_b200 import LSC
        import types
        dp = types.SimpleNamespace(rparticles=r, q_array=q, E=E_GEV)     # apply() duck-types these three
        if True:
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
        sec = cpu_kicks(ns, mesh, 3, "reference", 1)
        line["cpu_baseline"] = {"value": ns / sec, "unit": UNIT, "cores": 1, "kind": "port",
                                "sample": f"median of 3 kicks (1 warm-up) of {ns} particles on {mesh}^3, oracle port "
                                          f"in the reference's own configuration (numpy.fft on the (2n-1)^3 box, "
                                          f"single thread; host has {os.cpu_count()} cores)"}
    if line is not None:
        print(json.dumps(line), flush=True)
    if world > 1:
        if sharded is not None:
            sharded.finalize()           # graphs that captured NCCL work must be destroyed before the communicator
        torch.cuda.synchronize()
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_native(a)
