#!/usr/bin/env python
"""bench.py -- ISF+S(q) evaluations/s on BASELINE.json's C2 workload (N=256 He-4, M=170, 64 q), 3-D.

    python bench.py --gpus N --steps K --warmup W            (N>1: launched under torch.distributed.run)
    python bench.py --impl reference ...                     (the CPU restatement of the reference estimators)

A step = one pass of the hot path (rho_q build, tau-correlation, bin accumulation) over one batch of B synthetic
walker configurations per GPU.  `value` = configurations evaluated per second over all ranks with beads resident
in HBM; `e2e` = the same through the C ABI from pinned HOST buffers (H2D of every batch and D2H of the bin inside
the timed region).  One JSON line on stdout (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from pimc_b200 import synth  # noqa: E402

METRIC = "ISF+S(q) evaluations/sec (N=256 He-4, M=170, 64 q)"
UNIT = "evaluations/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=64, help="walker configurations per GPU per step")
    ap.add_argument("--workload", default="C2", choices=sorted(synth.SHAPES))
    ap.add_argument("--rho-mode", type=int, default=-1, help="-1 library default, 0 generic sincos, 1 lattice recurrence")
    ap.add_argument("--corr-mode", type=int, default=-1, help="-1 library default, 0 CUDA-core tau-correlation, 1 DMMA")
    ap.add_argument("--profile", default="rho", help="kernels timed with events INSIDE the timed region: all | rho | none")
    ap.add_argument("--profile-stride", type=int, default=8, help="bracket every n-th launch of a profiled kernel")
    ap.add_argument("--shard", default="config", choices=["config", "q"],
                    help="multi-GPU axis: independent walker configurations (weak scaling, one reduce) or q-vectors "
                         "(strong scaling, every rank sees every configuration, one all-gather)")
    ap.add_argument("--collective", default="torch", choices=["torch", "lib"],
                    help="who issues the one collective per bin: torch.distributed on the library's device buffer, or the "
                         "library's own NCCL entry points (pimcb_reduce_bins / pimcb_gather_bins_q)")
    ap.add_argument("--unique", type=int, default=16, help="distinct synthetic configurations generated per slot")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-numa", action="store_true", help="do not bind the rank to its GPU's local CPUs")
    ap.add_argument("--no-latency", action="store_true", help="skip the single-configuration latency leg")
    ap.add_argument("--no-ab", action="store_true", help="skip the generic-kernel A/B leg")
    ap.add_argument("--no-pair", action="store_true", help="skip the pair-potential secondary measurement")
    ap.add_argument("--peak-seconds", type=float, default=0.5, help="duration of the in-run FP64 DFMA peak measurement")
    ap.add_argument("--cpu-seconds", type=float, default=20.0, help="CPU work budget (core-seconds) of the cpu_baseline sample")
    return ap.parse_args()


def host_cores() -> int:
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return os.cpu_count() or 1


def workload_q(shape):
    if shape.name == "C3":
        n = np.stack(np.meshgrid(np.arange(-8, 9), np.arange(-8, 9), indexing="ij"), axis=-1).reshape(-1, 2)
        return (2.0 * np.pi / shape.side) * n          # max_int "8 8": odometer order, last dim fastest
    return synth.commensurate_q(shape.nq, shape.side)


def kernel_algorithmic_flops(shape, nq, plan):
    """Useful flop per evaluated configuration of the rho_q kernel that actually ran.
    generic: SURVEY 8d convention.  lattice (DMMA or CUDA-core): per bead ND*(1 + 40) for the base phases, 6 per extra
    power, 8 per (a>0,b>0) column product pair; per (group, bead) 2 flop per real sum K (8 in 3-D, 4 in 2-D, 2 in 1-D)."""
    if plan["path"] == 0:
        return algorithmic_flops(shape, nq)[0]
    nd = shape.ndim
    nmax = [plan["nmax_x"], plan["nmax_y"], plan["nmax_z"]][:nd]
    per_bead = nd * 41 + 6 * sum(max(n - 1, 0) for n in nmax)
    if nd == 3:
        per_bead += 8 * max(0, (plan["L_rows"] - 1) // 4)            # rough: one X*Y / X*conj(Y) pair per 4 L rows
    nk = {3: 8, 2: 4, 1: 2}[nd]
    return shape.N * shape.M * (per_bead + 2 * nk * plan["groups"])


def algorithmic_flops(shape, nq):
    """rho_q build and tau-correlation flop per evaluated configuration."""
    rho = nq * shape.N * shape.M * (2 * shape.ndim + 40 + 2)
    corr = nq * shape.M * (shape.M // 2 + 1) * 4
    return rho, corr


# --------------------------------------------------------------------------------------------------------------
# CPU arm: the reference's CPU estimators on the host cores (bounded sample, linear extrapolation by term count)
# --------------------------------------------------------------------------------------------------------------
class UpstreamCpu:
    """oracle/_ref/librefcpu<NDIM>d_fast.so: the reference's OWN accumulate() bodies (cut out of the upstream tree at build
    time, oracle/ref_cpu_extract.py + ref_cpu_shim.cpp) compiled with the reference's optimisation flags."""

    def __init__(self, ndim):
        import ctypes as C
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle", "_ref", f"librefcpu{ndim}d_fast.so")
        self.lib = C.CDLL(path)                       # OSError when it was not built: the caller falls back to the port
        dp, up = C.POINTER(C.c_double), C.POINTER(C.c_uint)
        self.lib.refcpu_ssf.argtypes = [dp, up, dp, C.c_int, C.c_int, C.c_int, dp, C.c_int, dp]
        self.lib.refcpu_isf.argtypes = [dp, dp, C.c_int, C.c_int, C.c_int, dp, C.c_int, dp]
        self.dp, self.up = dp, up

    def ssf(self, side, beads, N, q):
        M, Next, nd = beads.shape
        out = np.zeros(len(q))
        per = np.ones(nd, dtype=np.uint32)
        self.lib.refcpu_ssf(side.ctypes.data_as(self.dp), per.ctypes.data_as(self.up), beads.ctypes.data_as(self.dp), M, N, Next,
                            q.ctypes.data_as(self.dp), len(q), out.ctypes.data_as(self.dp))
        return out

    def isf(self, side, beads, N, q):
        M, Next, nd = beads.shape
        out = np.zeros((len(q), M))
        self.lib.refcpu_isf(beads.ctypes.data_as(self.dp), side.ctypes.data_as(self.dp), M, N, Next, q.ctypes.data_as(self.dp), len(q),
                            out.ctypes.data_as(self.dp))
        return out


class CpuArm:
    """Times the reference's CPU S(q) + F(q,tau) on all host threads.  kind = "reference": the upstream accumulate() bodies
    themselves (UpstreamCpu); the upstream F(q,tau) loop cannot be restricted to a tau range, so its bounded sample is the
    full (q, t0, tau, i, j) loop nest over the first Ms time slices of the configuration for one q per thread -- every term
    costs the same cos/sin quadruple -- extrapolated by the term count (nq/threads) * (M/Ms)^2; S(q) runs at full size on
    a subset of q.  kind = "port" (when the upstream library was not built): the oracle restatement, sampled by output
    element."""

    def __init__(self, shape, q, budget_core_s):
        self.shape, self.q = shape, np.ascontiguousarray(q, dtype=np.float64)
        self.cores = host_cores()
        self.beads = synth.gen_config(shape.N, shape.M, shape.ndim, shape.rho, shape.T)
        self.side = np.ascontiguousarray(shape.side, dtype=np.float64)
        c = self.cores
        try:
            self.up = UpstreamCpu(shape.ndim)
            self.kind = "reference"
        except OSError:
            self.up = None
            self.kind = "port"
            from oracle import oracle
            self.orc = oracle.get(fast=True)
        # per-term cost estimates (measured on the container's Xeon) only size the sample
        ssf_q_s = shape.M * shape.N * (shape.N - 1) / 2 * 40e-9
        n_ssf = 0.4 * budget_core_s / ssf_q_s
        self.n_ssf = int(min(len(q), max(1, round(n_ssf / c)) * c if n_ssf >= c else max(1, round(n_ssf))))
        if self.up:
            per_thread_s = 0.6 * budget_core_s / c
            ms = int((per_thread_s / (shape.N ** 2 * 25e-9)) ** 0.5)
            self.Ms = max(2, min(shape.M, ms - ms % 2))
            self.nq_isf = min(len(q), c)
            self.sub = np.ascontiguousarray(self.beads[:self.Ms])
        else:
            isf_elem_s = shape.M * shape.N**2 * 25e-9
            self.n_elem = int(min(len(q) * shape.M, max(1, round(0.6 * budget_core_s / isf_elem_s / c)) * c))   # whole rounds of threads

    def step(self):
        """One bounded sample; returns (extrapolated seconds per full evaluation, wall seconds of the sample)."""
        s, q = self.shape, self.q
        if not self.up:
            t0 = time.perf_counter()
            self.orc.isf_range(self.beads, s.N, q, 0, self.n_elem, nthreads=self.cores)
            t1 = time.perf_counter()
            self.orc.ssf(s.side, self.beads, s.N, q[:self.n_ssf], nthreads=min(self.cores, self.n_ssf))
            t2 = time.perf_counter()
            full = (t1 - t0) * (len(q) * s.M / self.n_elem) + (t2 - t1) * (len(q) / self.n_ssf)
            return full, t2 - t0
        from concurrent.futures import ThreadPoolExecutor          # ctypes calls release the GIL: one upstream loop per host thread
        with ThreadPoolExecutor(self.cores) as pool:
            t0 = time.perf_counter()
            list(pool.map(lambda k: self.up.isf(self.side, self.sub, s.N, np.ascontiguousarray(q[k:k + 1])), range(self.nq_isf)))
            t1 = time.perf_counter()
            nt = min(self.cores, self.n_ssf)
            chunks = [np.ascontiguousarray(q[:self.n_ssf][k::nt]) for k in range(nt)]
            list(pool.map(lambda qq: self.up.ssf(self.side, self.beads, s.N, qq), chunks))
            t2 = time.perf_counter()
        rounds = -(-len(q) // self.cores)                            # q-vectors per thread for the whole q-set
        full = (t1 - t0) * rounds * (s.M / self.Ms) ** 2 + (t2 - t1) * (len(q) / self.n_ssf)
        return full, t2 - t0

    def one_core(self):
        """Extrapolated seconds per full evaluation on ONE host core (the reference is single-threaded): the same bounded
        sample, one q through the F(q,tau) loop nest and a few q through S(q), on the calling thread."""
        s, q = self.shape, self.q
        n1 = max(1, self.n_ssf // self.cores)
        t0 = time.perf_counter()
        if self.up:
            self.up.isf(self.side, self.sub, s.N, np.ascontiguousarray(q[:1]))
            t1 = time.perf_counter()
            self.up.ssf(self.side, self.beads, s.N, np.ascontiguousarray(q[:n1]))
            t2 = time.perf_counter()
            return (t1 - t0) * len(q) * (s.M / self.Ms) ** 2 + (t2 - t1) * (len(q) / n1)
        ne = max(1, self.n_elem // self.cores)
        self.orc.isf_range(self.beads, s.N, q, 0, ne, nthreads=1)
        t1 = time.perf_counter()
        self.orc.ssf(s.side, self.beads, s.N, q[:n1], nthreads=1)
        t2 = time.perf_counter()
        return (t1 - t0) * (len(q) * s.M / ne) + (t2 - t1) * (len(q) / n1)

    def sample_text(self):
        s = self.shape
        if self.up:
            return (f"upstream accumulate() bodies (oracle/_ref/librefcpu{s.ndim}d_fast.so), 1 configuration, {self.cores} threads: F(q,tau) full "
                    f"loop nest over the first {self.Ms} of {s.M} slices for {self.nq_isf} q (one per thread), S(q) at full size for "
                    f"{self.n_ssf} of {len(self.q)} q; extrapolated by term count to {len(self.q)} q x {s.M}^2 slice pairs")
        return (f"1 configuration: {self.n_elem} of {len(self.q) * s.M} F(q,tau) elements (direct O(M N^2) loop each) + "
                f"S(q) for {self.n_ssf} of {len(self.q)} q, {self.cores} threads, extrapolated linearly to the full q-set")


def run_reference(args, shape, q):
    """--impl reference: the reference's own CPU estimators (the upstream accumulate() bodies compiled into oracle/_ref,
    else the oracle port) with all host threads.  Under torchrun only rank 0 works."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    total_steps = max(1, args.steps + args.warmup)
    budget = min(args.cpu_seconds, 150.0 * host_cores() / total_steps)       # keep the whole run within minutes
    arm = CpuArm(shape, q, budget)
    for _ in range(args.warmup):
        arm.step()
    evals, wall = [], []
    for _ in range(args.steps):
        f, w = arm.step()
        evals.append(f)
        wall.append(w)
    value = 1.0 / float(np.mean(evals))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean(wall)), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{shape.name}: N={shape.N} M={shape.M} nq={len(q)} ndim={shape.ndim}, He-4 SVP density, "
                               f"commensurate q; reference CPU estimators ({'upstream code' if arm.kind == 'reference' else 'oracle port'}) "
                               f"on a bounded sample per step",
                   "parallelism": f"{arm.cores} host threads, q-vectors / F(q,tau) elements split over threads"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": arm.cores, "kind": arm.kind, "sample": arm.sample_text()},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# --------------------------------------------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
               0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        while not self._stop_evt.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                try:
                    r = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.REASONS.items():
                    if r & bit and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.004)

    def stop(self):
        self._stop_evt.set()
        if self.is_alive():
            self.join()
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def bind_to_gpu_numa(local: int):
    """Pin this rank's threads (and so the first-touch placement of its pinned staging buffers) to the CPUs NVML
    reports as local to its GPU -- what `numactl` per rank would do.  With 8 ranks on a two-socket host, buffers on
    the wrong socket halve the aggregate host-to-device rate.  Returns the CPU list or None."""
    try:
        import pynvml
        import torch
        pynvml.nvmlInit()
        p = torch.cuda.get_device_properties(local)
        try:
            bus = "%08x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
            h = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
        except Exception:
            h = pynvml.nvmlDeviceGetHandleByIndex(local)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, ((os.cpu_count() or 64) + 63) // 64)
        cpus = [64 * i + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1]
        allowed = os.sched_getaffinity(0)
        cpus = [c for c in cpus if c in allowed]
        if cpus:
            os.sched_setaffinity(0, cpus)
            return cpus
    except Exception:
        pass
    return None


# --------------------------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------------------------
def run_ours(args, shape, q):
    import torch
    import torch.distributed as dist

    from pimc_b200 import api

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    torch.cuda.set_device(local)
    all_cpus = os.sched_getaffinity(0)
    numa_cpus = bind_to_gpu_numa(local) if not args.no_numa else None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    from pimc_b200 import multi
    B, K, W = args.batch, args.steps, args.warmup
    q_all = q
    if args.shard == "q" and world > 1:
        lo, hi = multi.shard_range(len(q_all), world, rank)
        q = q_all[lo:hi]
    ctx = api.Context(local, shape.ndim)
    ctx.set_box(shape.side)
    ctx.set_qvecs(q)
    if args.rho_mode >= 0:
        ctx.set_rho_mode(args.rho_mode)
    if args.corr_mode >= 0:
        ctx.set_corr_mode(args.corr_mode)
    nq = len(q)
    lib_coll = world > 1 and args.collective == "lib"
    if lib_coll:
        uid = [api.Context.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)                    # plumbing: ship the 128-byte NCCL id to the ranks
        ctx.comm_init(world, rank, uid[0])

    # synthetic batches: `unique` distinct configurations per slot, cycled to B, different seeds per rank and slot
    nslots = ctx.num_slots()
    batch_bytes = B * shape.M * (shape.N + 3) * shape.ndim * 8
    use_slots = nslots if nslots * batch_bytes > 140e6 else nslots        # rotate all slots; total footprint below
    pinned = []
    for sl in range(use_slots):
        seed_rank = 0 if args.shard == "q" else rank          # q-sharding: every rank measures the same walkers
        uniq = synth.gen_batch(shape, min(args.unique, B), first=1000 * seed_rank + 100 * sl)
        pa = api.PinnedArray((B,) + uniq.shape[1:])
        for b in range(B):
            pa.array[b] = uniq[b % len(uniq)]
        pinned.append(pa)
    for sl, pa in enumerate(pinned):
        ctx.stage(pa.array, shape.N, slot=sl)
    ctx.sync()
    footprint_mb = use_slots * B * shape.M * 16 * ((shape.N + 15) // 16) * shape.ndim * 8 / 1e6

    ext = torch.cuda.ExternalStream(ctx.stream(), device=torch.device("cuda", local))
    peak_tflops = ctx.fp64_peak_tflops(args.peak_seconds)

    def device_step(k):
        ctx.select_slot(k % use_slots)
        ctx.measure()

    def bin_collective():
        # the one collective of the path, once per bin (not per step), on the library's own device buffer
        if lib_coll:
            if args.shard == "config":
                ctx.reduce_bins(0)
            else:
                g_ssf, _ = ctx.gather_bins_q(multi.shard_sizes(len(q_all), world))
                assert g_ssf.shape == (len(q_all),)
            return
        bins = multi.bins_tensor(ctx, torch.device("cuda", local))
        with torch.cuda.stream(ext):
            if args.shard == "config":
                dist.reduce(bins, dst=0, op=dist.ReduceOp.SUM)            # sum of per-GPU bins over NVLink
            else:
                loc = torch.cat([bins[:nq, None], bins[nq:].view(nq, shape.M)], dim=1)
                gathered = multi.gather_q_shards(loc, len(q_all))         # concatenate the q-shards
                assert gathered.shape == (len(q_all), 1 + shape.M)

    # ---- value: device-resident --------------------------------------------------------------------------
    # events bracket the dominant kernel (rho_q build) on every launch of the timed region; the two small kernels are
    # timed in a separate pass below (an event between two kernels breaks their back-to-back staging)
    ctx.set_profiling({"all": True, "none": False}.get(args.profile, ["rho"]))
    ctx.set_profiling_stride(args.profile_stride)        # every 8th launch: an event pair costs ~3 us of stream time
    for k in range(W):
        device_step(k)
    if world > 1:
        bin_collective()       # warm-up of the collective too: NCCL connects its channels lazily on first use
    ctx.kernel_times(reset=True)
    ctx.reset_bins()
    launches0 = ctx.launch_count()
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(ext)
    for k in range(K):
        device_step(k)
    if world > 1:
        bin_collective()
    e1.record(ext)
    barrier()
    clocks = sampler.stop()
    ms = torch.tensor([e0.elapsed_time(e1)], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    launches = ctx.launch_count() - launches0
    ktimes = ctx.kernel_times(reset=True)
    ctx.set_profiling(False)
    ctx.set_profiling_stride(1)
    evals_per_step = world * B if args.shard == "config" else B    # q-sharding: all ranks work on the same B walkers
    value = evals_per_step * K / (ms_total * 1e-3)
    _, _, n_acc = ctx.read_bins()
    expect_acc = B * K * (world if (lib_coll and args.shard == "config" and rank == 0) else 1)   # the library's reduce sums the counts too
    assert n_acc == expect_acc, (n_acc, expect_acc)
    if args.profile != "all":                      # untimed-region pass with every kernel bracketed: corr / bins durations
        ctx.set_profiling(True)
        for k in range(max(10, min(K, 50))):
            device_step(k)
        kall = ctx.kernel_times(reset=True)
        for name in ("corr", "bins"):
            ktimes[name] = kall[name]
        if args.profile == "none":
            ktimes["rho"] = kall["rho"]
        ctx.reset_bins()
        ctx.set_profiling(False)

    # ---- e2e: host AoS buffers through the C ABI, H2D + D2H inside the timed region -------------------------
    e2e = None
    if not args.no_e2e:
        Ke = max(4, min(K, 60))
        ctx.reset_bins()
        ctx.stage(pinned[0].array, shape.N)
        for k in range(3):                               # warm-up of the pipelined loop
            ctx.measure()
            ctx.stage_async(pinned[(k + 1) % use_slots].array, shape.N)
            ctx.read_bins()
        barrier()
        t0 = time.perf_counter()
        for k in range(Ke):
            ctx.measure()                                               # async: kernels on batch k
            ctx.stage_async(pinned[(k + 1) % use_slots].array, shape.N) # H2D of batch k+1 overlaps them, DMAs back to back
            ssf_bin, isf_bin, _ = ctx.read_bins()                       # D2H of the step's result (syncs)
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        e2e = {"value": evals_per_step * Ke / float(dt.item()), "unit": UNIT,
               "h2d_bytes_per_step": int(batch_bytes), "d2h_bytes_per_step": int((nq + nq * shape.M) * 8),
               "steps": Ke, "path": "pimcb_stage_batch_async(pinned host AoS) + pimcb_measure + pimcb_read_bins, double-buffered"}

    # ---- latency: ONE configuration through the synchronous ABI calls an estimator's accumulate() makes ----------
    latency = None
    if not args.no_e2e and not args.no_latency:
        one = np.ascontiguousarray(pinned[0].array[0])                   # pageable host copy, like Path::beads
        ssf1, isf1 = ctx.stage(one, shape.N).ssf_isf()
        for _ in range(5):
            ctx.stage(one, shape.N).ssf_isf()
        nlat = 50
        t0 = time.perf_counter()
        for _ in range(nlat):
            ctx.stage(one, shape.N).ssf_isf()                             # pimcb_stage_beads + pimcb_ssf_isf (H2D, kernels, D2H, sync)
        lat = (time.perf_counter() - t0) / nlat
        one_locked = pinned[0].array[0]                                   # page-locked, as B200Session registers Path::beads
        for _ in range(5):
            ctx.ssf_isf_beads(one_locked, shape.N)
        t0 = time.perf_counter()
        for _ in range(nlat):
            ctx.ssf_isf_beads(one_locked, shape.N)                        # one ABI call: DMA, transpose, kernels, read-back, one sync
        lat_locked = (time.perf_counter() - t0) / nlat
        latency = {"single_configuration_us": lat_locked * 1e6, "evaluations_per_s": 1.0 / lat_locked, "calls": nlat,
                   "path": "pimcb_ssf_isf_beads(page-locked host AoS), what B200Session calls: one synchronous ABI call per walker"
                           + ("" if os.environ.get("PIMCB_GRAPH", "1") == "0" else
                              "; H2D -> transpose -> rho_q -> tau-correlation -> D2H replayed as one CUDA graph"),
                   "pageable_source_us": lat * 1e6}
        for sl, pa in enumerate(pinned):               # the single-walker calls cycled through the slots: restore the batches
            ctx.stage(pa.array, shape.N, slot=sl)
        ctx.sync()
        ctx.reset_bins()

    # ---- roofline of the dominant kernel (rho_q build) -------------------------------------------------------
    rho_flop, corr_flop = algorithmic_flops(shape, nq)
    rho_ms, rho_n = ktimes["rho"]
    corr_ms, corr_n = ktimes["corr"]
    bins_ms, _ = ktimes["bins"]
    rho_avg_s = rho_ms * 1e-3 / max(1, rho_n)
    plan = ctx.rho_plan_info()
    kernel_flop = kernel_algorithmic_flops(shape, nq, plan)          # useful flop of the kernel that actually ran
    achieved = B * kernel_flop / rho_avg_s / 1e12 if rho_n else None
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get(plan["path_name"] + "_dram_bytes_per_launch")
        except Exception:
            traffic = None
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    alg_bytes = B * (8 * shape.ndim * shape.N * shape.M + 2 * 8 * nq * shape.M)
    roofline = {
        "kernel": plan["path_name"],
        "bound": "fp64" if plan["path"] != 1 else "fp64 (DFMA phase A + DMMA tensor phase B share the FP64 units)",
        "achieved": achieved, "peak": peak_tflops, "unit": "TFLOP/s",
        "frac": (achieved / peak_tflops) if achieved else None, "traffic": traffic,
        "peak_source": "measured in this run: register-resident DFMA chains on all SMs (pimcb_measure_fp64_peak); "
                       "MEASURED_PEAKS.json has no FP64 figure (DMMA m8n8k4 measures 37.0 TFLOP/s, tools/micro/dmma_peak.cu)",
        "flop_per_launch": B * kernel_flop, "flop_basis": "useful flop of the kernel that ran (DESIGN.md section 5)",
        "avg_launch_ms": rho_avg_s * 1e3, "launches_timed": rho_n,
        "timing": f"CUDA events around every {args.profile_stride}-th launch of this kernel inside the timed region ({rho_n} of {K} steps)",
        "share_of_step": rho_avg_s * 1e3 / max(1e-12, rho_avg_s * 1e3 + corr_ms / max(1, corr_n) + bins_ms / max(1, ktimes["bins"][1])),
        "bins_kernel": {"avg_launch_ms": bins_ms / max(1, ktimes["bins"][1])},
        "plan": plan,
        # the same launch expressed in SURVEY.md section 8d's convention (one 40-flop sincos per (q, bead)): what a
        # generic kernel would have to sustain to finish in the same time
        "survey_convention": {"flop_per_launch": B * rho_flop, "equivalent_tflops": B * rho_flop / rho_avg_s / 1e12 if rho_n else None},
        "hbm": {"achieved_gbs": alg_bytes / rho_avg_s / 1e9 if rho_n else None, "peak_gbs": hbm_peak,
                "peak_source": "MEASURED_PEAKS.json" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"},
        "corr_kernel": {"avg_launch_ms": corr_ms / max(1, corr_n), "achieved_tflops": B * corr_flop / (corr_ms * 1e-3 / max(1, corr_n)) / 1e12 if corr_n else None},
    }
    # ---- A/B: the generic kernel (one sincos per (q, bead)) on the same batches, SURVEY 8d flop convention --------
    if plan["path"] != 0 and not args.no_ab:
        ctx.set_rho_mode(0)
        ctx.set_profiling(True)
        for k in range(3):
            device_step(k)
        ctx.kernel_times(reset=True)
        nab = max(5, min(K, 40))
        for k in range(nab):
            device_step(k)
        kt0 = ctx.kernel_times(reset=True)
        ctx.set_profiling(False)
        ctx.set_rho_mode(1 if args.rho_mode < 0 else args.rho_mode)
        g_s = kt0["rho"][0] * 1e-3 / max(1, kt0["rho"][1])
        g_ach = B * rho_flop / g_s / 1e12
        roofline["generic_kernel_ab"] = {"kernel": "rho_generic_kernel", "avg_launch_ms": g_s * 1e3, "launches_timed": kt0["rho"][1],
                                         "flop_per_launch": B * rho_flop, "achieved": g_ach, "peak": peak_tflops,
                                         "frac": g_ach / peak_tflops, "bound": "fp64",
                                         "note": "not part of the timed region; same inputs, same outputs to rounding"}
        ctx.reset_bins()

    # ---- secondary metric: per-slice pair-potential sums (Vint, gradVSquared on odd slices, sepHist) -------------
    pair = None
    if shape.ndim == 3 and not args.no_pair:
        import math
        max_sep = math.sqrt(sum((L / 2.0) ** 2 for L in shape.side))
        Vt, dVt, drt = synth.aziz_table_numpy(max_sep)
        ctx.set_pair_table(Vt, dVt, drt)
        ctx.select_slot(0)
        dSep = 0.5 * math.sqrt(3.0) * shape.side[2] / 50.0
        ctx.set_profiling(True)
        for _ in range(2):
            ctx.pair_sums(dSep, want_f2=True, want_hist=True, f2_parity=1)
        ctx.kernel_times(reset=True)
        npair = 5
        for k in range(npair):
            ctx.select_slot(k % use_slots)
            ctx.pair_sums(dSep, want_f2=True, want_hist=True, f2_parity=1)
        pms, pn = ctx.kernel_times(reset=True)["pair"]
        ctx.set_profiling(False)
        p_s = pms * 1e-3 / max(1, pn)
        pairs = shape.N * (shape.N - 1) // 2               # pairs per slice
        # gsf action: even slices read V once per pair; odd slices read V and dV/dr once per pair in the symmetric kernel
        # (pair_sym_kernel), or walk the full j != i loop in the both-ends kernel (2 visits per pair, each reads dV/dr,
        # half of them also read V: 3 table reads per pair)
        pair_sym = os.environ.get("PIMCB_PAIR_SYM", "1") != "0" and shape.N <= 1024      # every pair once on the force slices
        gathers = B * ((shape.M - shape.M // 2) * pairs + (shape.M // 2) * (2 if pair_sym else 3) * pairs)
        pair = {"metric": "pair-potential action sums/s (Vint[M] + gradVSquared[odd slices] + sepHist[M][50] per configuration)",
                "kernel": "pair_sym_kernel (force slices: every pair once)" if pair_sym else "pair_kernel (force slices: both ends)",
                "value": B / p_s, "unit": "configurations/s", "avg_launch_ms": p_s * 1e3, "launches_timed": pn,
                "table_entries": len(Vt), "table_mb": 2 * 8 * len(Vt) / 1e6,
                "gathers_per_launch": gathers, "gather_rate_g_per_s": gathers / p_s / 1e9,
                "sector_traffic_tbs": 32.0 * gathers / p_s / 1e12,
                "bound": "L2/HBM gather: 8-byte table reads move 32-byte sectors; the 106 MB of tables exceed what one L2 "
                         "partition keeps, ncu shows 4.5 GB of DRAM reads per launch for the symmetric kernel, 6.4 GB for the "
                         "both-ends kernel (profiles/traffic.json)"}

    # ---- secondary: virial slice sums (rDOTgradU / deltaDOTgradU, gsf: T-matrix terms on odd slices) ---------------
    if pair is not None:
        d2Vt = np.gradient(dVt, drt)                   # timing only: central differences of the dV/dr table
        ctx.set_pair_table_d2(d2Vt)
        ctx.select_slot(0)
        delta = 0.01 * pinned[0].array                 # any per-bead vectors in the beads' AoS shape
        ctx.set_profiling(True)
        ctx.virial_sums(delta, t2_parity=1)
        ctx.kernel_times(reset=True)
        nvir = 3
        for k in range(nvir):
            ctx.virial_sums(delta, t2_parity=1)
        vms, vn = ctx.kernel_times(reset=True)["virial"]
        ctx.set_profiling(False)
        v_s = vms * 1e-3 / max(1, vn)
        vir_sym = os.environ.get("PIMCB_VIRIAL_SYM", "1") != "0" and shape.N <= 1024
        vg = B * shape.N * (shape.N - 1) * (shape.M + shape.M // 2) // (2 if vir_sym else 1)   # dV/dr on every slice, d2V/dr2 on odd ones
        pair["virial_sums"] = {"metric": "virial slice sums/s (4 sums per slice; gsf action, window deltas from the host)",
                               "kernel": "virial_sym_kernel (every pair once)" if vir_sym else "virial_kernel (both ends)",
                               "value": B / v_s, "unit": "configurations/s", "avg_launch_ms": v_s * 1e3, "launches_timed": vn,
                               "gathers_per_launch": vg, "gather_rate_g_per_s": vg / v_s / 1e9,
                               "bound": "L2/HBM gather (same tables as the pair kernel + d2V/dr2)"}

    # ---- secondary: direct minimum-image S(q) for non-commensurate (`float`) wave-vectors --------
    direct = None
    if not args.no_pair:
        nqd = 8
        qd = synth.float_q(nqd, shape.ndim)
        dctx = api.Context(local, shape.ndim)
        dctx.set_box(shape.side)
        dctx.set_qvecs(qd)
        nb = min(B, 16)
        dctx.stage(pinned[0].array[:nb], shape.N)
        dctx.measure()
        dctx.set_profiling(True)
        nrep = 5
        for _ in range(nrep):
            dctx.measure()
        kt = dctx.kernel_times(reset=True)
        dctx.set_profiling(False)
        dkey = [k for k in kt if "direct" in k][0]
        d_s = kt[dkey][0] * 1e-3 / max(1, kt[dkey][1])
        pairs_q = nb * nqd * shape.M * shape.N * (shape.N - 1) // 2
        dflop = pairs_q * (shape.ndim * 8 + 20 + 1)               # 45 flop per (pair, q) in 3-D
        direct = {"metric": "direct min-image S(q), non-commensurate q (ssf_direct_kernel)", "nq": nqd, "configurations": nb,
                  "avg_launch_ms": d_s * 1e3, "launches_timed": kt[dkey][1], "pair_q_terms_per_launch": pairs_q,
                  "achieved_tflops": dflop / d_s / 1e12, "peak_tflops": peak_tflops, "frac": dflop / d_s / 1e12 / peak_tflops,
                  "terms_per_s": pairs_q / d_s, "bound": "fp64"}
        dctx.close()

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak" if args.shard == "config" else "strong",
        "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{shape.name}: N={shape.N} M={shape.M} nq={nq} ndim={shape.ndim}, He-4 SVP density, "
                               f"commensurate q, {B} walker configurations per GPU per step",
                   "batch_per_gpu": B,
                   "parallelism": (f"walker-configuration sharding x{world}, one NCCL reduce of the bin" if args.shard == "config"
                                   else f"q-vector sharding x{world} ({nq} of {len(q_all)} q per GPU), one NCCL all-gather of the bin"),
                   "l2": f"{use_slots} resident batches rotated ({footprint_mb:.0f} MB > 126 MB L2)",
                   "rho_mode": args.rho_mode, "corr_mode": args.corr_mode,
                   "collective": ("none (one GPU)" if world == 1 else
                                  ("library NCCL (pimcb_reduce_bins / pimcb_gather_bins_q)" if lib_coll else "torch.distributed NCCL on the library's bin")),
                   "host_binding": (f"rank bound to {len(numa_cpus)} GPU-local CPUs (NVML affinity)" if numa_cpus else "none")},
        "clocks": clocks, "e2e": e2e, "latency": latency, "gpu_launches": int(launches), "roofline": roofline, "pair_sums": pair, "ssf_direct": direct,
    }

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        os.sched_setaffinity(0, all_cpus)             # the CPU arm gets every host core again
        arm = CpuArm(shape, q, args.cpu_seconds)
        full_s, _ = arm.step()
        line["cpu_baseline"] = {"value": 1.0 / full_s, "unit": UNIT, "cores": arm.cores, "kind": arm.kind,
                                "sample": arm.sample_text(), "one_core_value": 1.0 / arm.one_core()}
    else:
        line["cpu_baseline"] = None
    if rank == 0:
        emit(line)
    for pa in pinned:
        pa.free()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


_REAL_STDOUT = None


def claim_stdout():
    """Everything any library prints on fd 1 (NCCL's version banner, torchrun notices) goes to stderr; the JSON line is
    written to the original stdout by emit()."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line: dict):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    args = parse_args()
    shape = synth.SHAPES[args.workload]
    q = workload_q(shape)
    if args.impl == "reference":
        run_reference(args, shape, q)
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 and world == 1:
        # convenience: re-launch under torch.distributed.run, one rank per GPU
        import subprocess
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", os.environ.get("MASTER_PORT", "29511"), os.path.abspath(__file__)] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    claim_stdout()
    run_ours(args, shape, q)


if __name__ == "__main__":
    main()
