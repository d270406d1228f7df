#!/usr/bin/env python
"""bench.py -- ISF+S(q) evaluations/s on BASELINE.json's C2 workload (N=256 He-4, M=170, 64 q), 3-D.

    python bench.py --gpus N --steps K --warmup W            (N>1: launched under torch.distributed.run)
    python bench.py --impl reference ...                     (the CPU restatement of the reference estimators)

A step = ONE OUTPUT BIN of the hot path: `--batches-per-step` (16) batches of B (512) synthetic walker configurations per GPU
go through rho_q build + tau-correlation + bin accumulation, then the bin is folded and -- on more than one GPU --
exchanged once (the library's own NCCL reduce / all-gather), exactly what EstimatorBase::output does per bin.
`value` = configurations evaluated per second over all ranks with beads resident in HBM; `e2e` = the same through the
C ABI from pinned HOST buffers (H2D of every batch and D2H of every bin inside the timed region), with the bare
H2D rate of the same buffers measured in the same run as its ceiling.  One JSON line on stdout (rank 0).
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

METRIC = "ISF+S(q) evaluations/sec (N=256 He-4, M=170, 64 q)"       # BASELINE.json's metric, quoted on C2


def metric_name(shape, nq):
    if shape.name == "C2":
        return METRIC
    return f"ISF+S(q) evaluations/sec ({shape.name}: N={shape.N} He-4, M={shape.M}, {nq} q, {shape.ndim}-D)"
UNIT = "evaluations/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=0,
                    help="0 = 512 (64 for C4, whose configurations are 7.5x larger); walker configurations per GPU per batch (one launch pair); "
                         "64 -> 256 -> 512 per launch is worth 16 %% + 4 %% (profiles/r02p_corr_occ_and_batch_sweep.txt: fewer ragged last "
                         "waves of the persistent rho kernel, more waves of the tau-correlation)")
    ap.add_argument("--batches-per-step", type=int, default=16,
                    help="batches accumulated into one output bin = one step (16 x 512 = 8192 evaluations per GPU)")
    ap.add_argument("--exchange", default="pipelined", choices=["pipelined", "inline"],
                    help="walker sharding on > 1 GPU with --collective lib: pimcb_reduce_bins_begin/_end (the bin's snapshot is reduced "
                         "on the library's communication stream while the next bin is measured) or pimcb_reduce_bins on the compute stream")
    ap.add_argument("--workload", default="C2", choices=sorted(synth.SHAPES))
    ap.add_argument("--rho-mode", type=int, default=-1, help="-1 library default, 0 generic sincos, 1 lattice recurrence")
    ap.add_argument("--corr-mode", type=int, default=-1, help="-1 library default, 0 CUDA-core tau-correlation, 1 DMMA")
    ap.add_argument("--profile", default="rho", help="kernels timed with events INSIDE the timed region: all | rho | none")
    ap.add_argument("--profile-stride", type=int, default=8, help="bracket every n-th launch of a profiled kernel")
    ap.add_argument("--shard", default="config", choices=["config", "q"],
                    help="multi-GPU axis: independent walker configurations (weak scaling, one reduce) or q-vectors "
                         "(strong scaling, every rank sees every configuration, one all-gather)")
    ap.add_argument("--collective", default="lib", choices=["torch", "lib"],
                    help="who issues the one collective per bin: torch.distributed on the library's device buffer, or the "
                         "library's own NCCL entry points (pimcb_reduce_bins / pimcb_gather_bins_q)")
    ap.add_argument("--total-batch", type=int, default=0,
                    help="STRONG scaling over walkers: this many configurations per batch in total, split over the GPUs "
                         "(overrides --batch; scaling is reported as strong)")
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

    def one_core(self, repeats=3):
        """Extrapolated seconds per full evaluation on ONE host core (the reference is single-threaded): the same bounded
        sample, one q through the F(q,tau) loop nest and a few q through S(q), on the calling thread; median of `repeats`
        (a single shot came out 3x slow once on a shared host: r02y)."""
        return float(np.median([self._one_core_once() for _ in range(max(1, repeats))]))

    def _one_core_once(self):
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
                    f"{self.n_ssf} of {len(self.q)} q; extrapolated by term count to {len(self.q)} q x {s.M}^2 slice pairs; median over the timed steps")
        return (f"1 configuration: {self.n_elem} of {len(self.q) * s.M} F(q,tau) elements (direct O(M N^2) loop each) + "
                f"S(q) for {self.n_ssf} of {len(self.q)} q, {self.cores} threads, extrapolated linearly to the full q-set; median over the timed steps")


def workload_text(shape, nq):
    return (f"{shape.name}: N={shape.N} M={shape.M} nq={nq} ndim={shape.ndim}, He-4 SVP density, commensurate q; one evaluation = "
            f"S(q)[{nq}] + F(q,tau)[{nq}][{shape.M}] of one walker configuration")


def shared_config(shape, nq):
    """`config` is the WORKLOAD and is identical in both arms (the driver compares it); how each arm runs it is under `run`."""
    return {"workload": workload_text(shape, nq)}


FULL_SIZE_RECORD = {"value": 1.0 / 261.0, "unit": UNIT, "cores": 8,
                    "what": "ONE complete C2 evaluation through the upstream CPU accumulate() bodies, no sampling: 259.7 s F(q,tau) + "
                            "1.3 s S(q) on the 8 threads of the build container (tests/golden/make_upstream_c2_full.py -> "
                            "tests/golden/upstream_c2_full.npz; the CUDA path reproduces all 64 + 10 880 values to 1e-10)"}


def time_cpu_arm(arm, steps, warmup):
    for _ in range(warmup):
        arm.step()
    evals, wall = [], []
    for _ in range(steps):
        f, w = arm.step()
        evals.append(f)
        wall.append(w)
    # median over the steps: the sample is statically partitioned (one q per thread), so one descheduled thread stretches
    # a whole step -- the mean of a handful of steps moved by 10-20 % between two processes on the same box
    return 1.0 / float(np.median(evals)), float(np.median(wall))


def run_reference(args, shape, q):
    """--impl reference: the reference's own CPU estimators (the upstream accumulate() bodies compiled into oracle/_ref,
    else the oracle port) with all host threads, on the same workload and with the same bounded sample per step as the
    cpu_baseline leg of the GPU arm.  Under torchrun only rank 0 works."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    total_steps = max(1, args.steps + args.warmup)
    budget = min(args.cpu_seconds, 240.0 * host_cores() / total_steps)       # keep the whole run within minutes
    arm = CpuArm(shape, q, budget)
    value, wall = time_cpu_arm(arm, args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": metric_name(shape, len(q)), "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * wall, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": shared_config(shape, len(q)),
        "run": {"what": f"reference CPU estimators ({'upstream code' if arm.kind == 'reference' else 'oracle port'}), a bounded sample of "
                        "the evaluation per step, extrapolated by term count (see cpu_baseline.sample)",
                "parallelism": f"{arm.cores} host threads, q-vectors / F(q,tau) elements split over threads"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": arm.cores, "kind": arm.kind, "sample": arm.sample_text(),
                         "extrapolated": True, "full_size_record": FULL_SIZE_RECORD},
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
def upstream_gpu_ab(shape, q, beads_one):
    """the reference's shipped GPU kernels (src/estimator_gpu.cu, compiled UNMODIFIED
    for sm_100a into oracle/_ref/librefgpu3d*.so by `make -C oracle ref`) timed on the same box, same configuration, same q:
    one gpu_ssf launch + one gpu_isf launch per q per evaluation, with the header's default block size (256) and with
    the one upstream's CMake passes (1024)."""
    import ctypes as C
    if shape.ndim != 3:
        return None
    out = {}
    dp = C.POINTER(C.c_double)
    beads = np.ascontiguousarray(beads_one, dtype=np.float64)
    qq = np.ascontiguousarray(q, dtype=np.float64)
    M, Next, _ = beads.shape
    for tag, name in (("block256", "librefgpu3d.so"), ("block1024", "librefgpu3d_b1024.so")):
        path = os.path.join(ROOT, "oracle", "_ref", name)
        if not os.path.exists(path):
            continue
        try:
            lib = C.CDLL(path)
            lib.ref_gpu_bench.argtypes = [dp, C.c_int, C.c_int, C.c_int, dp, C.c_int, C.c_int, dp]
            ms = np.zeros(2)
            if lib.ref_gpu_bench(beads.ctypes.data_as(dp), M, shape.N, Next, qq.ctypes.data_as(dp), len(qq), 2, ms.ctypes.data_as(dp)) != 0:
                continue
            out[tag] = {"kernels_ms_per_evaluation": float(ms[0]), "kernels_evaluations_per_s": 1e3 / float(ms[0]),
                        "accumulate_pattern_ms_per_evaluation": float(ms[1]), "accumulate_pattern_evaluations_per_s": 1e3 / float(ms[1])}
        except (OSError, AttributeError):
            continue
    if not out:
        return None
    best = max(out.values(), key=lambda v: v["kernels_evaluations_per_s"])
    return {"what": "upstream gpu_ssf + gpu_isf (src/estimator_gpu.cu:66-165, 288-413) compiled unmodified for sm_100a, one configuration, "
                    "2 evaluations timed after one warm-up; kernels = CUDA events with beads resident, accumulate_pattern = the "
                    "upstream estimators' own per-measurement H2D + launches + sync + D2H (src/estimator.cpp:3833-3861, 4074-4100)",
            "variants": out, "best_kernels_evaluations_per_s": best["kernels_evaluations_per_s"],
            "best_accumulate_evaluations_per_s": max(v["accumulate_pattern_evaluations_per_s"] for v in out.values())}


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

    def allreduce(x, op):
        t = torch.tensor([x], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=op)
        return float(t.item())

    from pimc_b200 import multi
    B, K, W, P = args.batch, args.steps, args.warmup, args.batches_per_step
    if B <= 0:
        B = 64 if shape.name == "C4" else 512
    if args.total_batch > 0:
        if args.total_batch % world:
            raise SystemExit(f"--total-batch {args.total_batch} is not a multiple of {world} GPUs")
        B = args.total_batch // world
    if P <= 0:
        P = 16
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
    use_slots = nslots
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

    pipelined = lib_coll and args.shard == "config" and args.exchange == "pipelined"
    xchg = {"pending": False, "last": None}

    def exchange_collect():
        """Wait for the exchange started one bin ago (root: the global bin lands in host memory)."""
        if xchg["pending"]:
            xchg["last"] = ctx.reduce_bins_end()
            xchg["pending"] = False
        return xchg["last"]

    def bin_exchange():
        """The end of an output bin: fold the accumulators and -- on more than one GPU -- the ONE collective of the path."""
        if world == 1:
            ctx.bins_device_ptr()                                 # enqueues the fold of the persistent rows into the bin
            return
        if lib_coll:
            if pipelined:
                # pimcb_reduce_bins_begin / _end: the previous bin's row is collected, then this bin's snapshot starts its
                # reduce on the library's communication stream while the next bin is measured
                exchange_collect()
                ctx.reduce_bins_begin(0)
                xchg["pending"] = True
            elif args.shard == "config":
                ctx.reduce_bins(0, want_total=False)          # no host synchronisation: the count comes with read_bins
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

    def device_step(k, exchange=True):
        for j in range(P):
            ctx.select_slot((k * P + j) % use_slots)
            ctx.measure()
        if exchange:
            bin_exchange()
            ctx.reset_bins()

    # ---- value: device-resident --------------------------------------------------------------------------
    # events bracket the dominant kernel (rho_q build) on every `profile_stride`-th launch of the timed region; the small
    # kernels are timed in a separate pass below (an event between two kernels breaks their back-to-back staging)
    ctx.set_profiling({"all": True, "none": False}.get(args.profile, ["rho"]))
    ctx.set_profiling_stride(args.profile_stride)        # every 8th launch: an event pair costs ~3 us of stream time
    for k in range(W):
        device_step(k)
    # one bin checked for its bookkeeping before the timed region: every configuration of every rank is in it
    device_step(0, exchange=False)
    bin_exchange()
    if pipelined:
        _, _, n_acc = exchange_collect()
        if rank != 0:
            n_acc = ctx.read_bins()[2]
    else:
        _, _, n_acc = ctx.read_bins()
    expect_acc = B * P * (world if (lib_coll and args.shard == "config" and rank == 0) else 1)
    assert n_acc == expect_acc, (n_acc, expect_acc)
    ctx.reset_bins()
    ctx.kernel_times(reset=True)
    launches0 = ctx.launch_count()
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(ext)
    for k in range(K):
        device_step(k)
    exchange_collect()                      # the last bin's row: inside the timed region
    e1.record(ext)
    barrier()
    clocks = sampler.stop()
    ms_total = allreduce(e0.elapsed_time(e1), dist.ReduceOp.MAX if world > 1 else None)
    launches = ctx.launch_count() - launches0
    ktimes = ctx.kernel_times(reset=True)
    ctx.set_profiling(False)
    ctx.set_profiling_stride(1)
    evals_per_step = (world * B if args.shard == "config" else B) * P    # q-sharding: all ranks work on the same walkers
    value = evals_per_step * K / (ms_total * 1e-3)
    if args.profile != "all":                      # untimed-region pass with every kernel bracketed: corr / bins durations
        ctx.set_profiling(True)
        for k in range(max(2, min(K, 4))):
            device_step(k)
        kall = ctx.kernel_times(reset=True)
        for name in ("corr", "bins"):
            ktimes[name] = kall[name]
        if args.profile == "none":
            ktimes["rho"] = kall["rho"]
        ctx.reset_bins()
        ctx.set_profiling(False)

    # ---- e2e: host AoS buffers through the C ABI, H2D of every batch + D2H of every bin inside the timed region ---------
    e2e = None
    if not args.no_e2e:
        Ke = max(3, min(K, 8))
        ctx.reset_bins()
        ctx.stage(pinned[0].array, shape.N)

        def e2e_step(k):
            for j in range(P):
                ctx.measure()                                                         # async: kernels on the staged batch
                ctx.stage_async(pinned[(k * P + j + 1) % use_slots].array, shape.N)   # H2D of the next batch overlaps them
            if world > 1:
                bin_exchange()
            if not pipelined:
                ssf_bin, isf_bin, _ = ctx.read_bins()                                 # D2H of the bin (syncs)
            # (pipelined exchange: the root's D2H of the GLOBAL bin is part of pimcb_reduce_bins_end, one bin later)
            ctx.reset_bins()

        e2e_step(0)                                          # warm-up of the pipelined loop
        barrier()
        t0 = time.perf_counter()
        for k in range(Ke):
            e2e_step(k)
        exchange_collect()
        torch.cuda.synchronize()
        dt = allreduce(time.perf_counter() - t0, dist.ReduceOp.MAX if world > 1 else None)
        # in-run ceiling: the bare cudaMemcpyAsync rate of the same page-locked buffers, all ranks copying at the same time
        ctx.sync()
        barrier()
        ceil_rank = ctx.h2d_peak_gbs(pinned[0].array, reps=12)
        barrier()
        ceil_min = allreduce(ceil_rank, dist.ReduceOp.MIN if world > 1 else None)
        ceil_sum = allreduce(ceil_rank, dist.ReduceOp.SUM if world > 1 else None)
        h2d_rate_rank = batch_bytes * P * Ke / dt / 1e9
        e2e = {"value": evals_per_step * Ke / dt, "unit": UNIT,
               "h2d_bytes_per_step": int(batch_bytes * P), "d2h_bytes_per_step": int((nq + nq * shape.M) * 8),
               "steps": Ke, "h2d_gbs_per_gpu": h2d_rate_rank,
               "h2d_ceiling_gbs": ceil_min, "h2d_ceiling_gbs_all_gpus": ceil_sum,
               "frac_of_ceiling": h2d_rate_rank / ceil_min if ceil_min else None,
               "ceiling": "bare cudaMemcpyAsync of one batch buffer, 12 back to back, all ranks concurrently, same run "
                          "(pimcb_measure_h2d_peak); min over ranks",
               "path": f"per step: {P} x [pimcb_measure + pimcb_stage_batch_async(pinned host AoS)] + "
                       + (("pimcb_reduce_bins_end(previous bin: root's D2H) + pimcb_reduce_bins_begin + pimcb_reset_bins" if pipelined else
                           "the bin collective + pimcb_read_bins + pimcb_reset_bins") if world > 1 else "pimcb_read_bins + pimcb_reset_bins")
                       + ", staging pipelined one batch ahead"}

    # ---- latency: ONE configuration through the synchronous ABI calls an estimator's accumulate() makes ----------
    latency = None
    if not args.no_e2e and not args.no_latency:
        one = np.ascontiguousarray(pinned[0].array[0])                   # pageable host copy, like Path::beads
        ctx.stage(one, shape.N).ssf_isf()
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
            tj = json.load(open(tpath))
            key = plan["path_name"] + "_dram_bytes_per_launch"
            traffic = tj.get(key)
            per = (tj.get("configurations_per_launch") or {}).get(key)
            if traffic is not None and per and shape.name == "C2":
                traffic = traffic * B / per            # the capture's launch held `per` configurations, this run's holds B
        except Exception:
            traffic = None
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    sm_max = (clocks or {}).get("sm_max_mhz") or 1965.0
    nominal_tflops = 148 * 64 * 2 * sm_max * 1e6 / 1e12               # 148 SMs x 64 DFMA lanes x 2 flop x clock
    alg_bytes = B * (8 * shape.ndim * shape.N * shape.M + 2 * 8 * nq * shape.M)
    corr_avg_s = corr_ms * 1e-3 / max(1, corr_n)
    roofline = {
        "kernel": plan["path_name"],
        "bound": "fp64" if plan["path"] != 1 else "fp64 (DFMA phase A + DMMA tensor phase B share the FP64 units)",
        "achieved": achieved, "peak": peak_tflops, "unit": "TFLOP/s",
        "frac": (achieved / peak_tflops) if achieved else None, "traffic": traffic,
        "peak_source": "measured in this run: register-resident DFMA chains on all SMs (pimcb_measure_fp64_peak; SASS and ncu of that "
                       "kernel: profiles/r02_fp64_peak.md); MEASURED_PEAKS.json has no FP64 figure",
        "frac_vs_nominal": (achieved / nominal_tflops) if achieved else None, "nominal_peak": nominal_tflops,
        "nominal_peak_source": f"148 SMs x 64 FP64 FMA lanes x 2 x {sm_max:.0f} MHz",
        "frac_vs_dmma_probe": (achieved / 36.9) if achieved else None, "dmma_probe_peak": 36.9,
        "dmma_probe_source": "tools/micro/dmma_peak.cu (mma.sync.m8n8k4.f64 chains), profiles/r01*_dmma_peak / DESIGN.md section 5",
        "flop_per_launch": B * kernel_flop, "flop_basis": "useful flop of the kernel that ran (DESIGN.md section 5)",
        "avg_launch_ms": rho_avg_s * 1e3, "launches_timed": rho_n,
        "configurations_per_launch": B, "us_per_64_configurations": rho_avg_s * 1e6 * 64.0 / B,
        "timing": f"CUDA events around every {args.profile_stride}-th launch of this kernel inside the timed region ({rho_n} of {K * P} launches)",
        "share_of_step": rho_avg_s / max(1e-12, rho_avg_s + corr_avg_s + bins_ms * 1e-3 / max(1, ktimes["bins"][1])),
        "bins_kernel": {"avg_launch_ms": bins_ms / max(1, ktimes["bins"][1])},
        "plan": plan,
        # the same launch expressed in SURVEY.md section 8d's convention (one 40-flop sincos per (q, bead)): what a
        # generic kernel would have to sustain to finish in the same time
        "survey_convention": {"flop_per_launch": B * rho_flop, "equivalent_tflops": B * rho_flop / rho_avg_s / 1e12 if rho_n else None},
        "hbm": {"achieved_gbs": alg_bytes / rho_avg_s / 1e9 if rho_n else None, "peak_gbs": hbm_peak,
                "peak_source": "MEASURED_PEAKS.json" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"},
        "corr_kernel": {"avg_launch_ms": corr_avg_s * 1e3, "us_per_64_configurations": corr_avg_s * 1e6 * 64.0 / B, "achieved_tflops": B * corr_flop / corr_avg_s / 1e12 if corr_n else None},
    }
    # ---- A/B: the generic kernel (one sincos per (q, bead)) on the same batches, SURVEY 8d flop convention --------
    if plan["path"] != 0 and not args.no_ab:
        ctx.set_rho_mode(0)
        ctx.set_profiling(True)
        device_step(0)
        ctx.kernel_times(reset=True)
        for k in range(2):
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
        pair = pair_legs(args, ctx, shape, pinned, use_slots, B)

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
        "metric": metric_name(shape, len(q_all)), "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_total / K, "higher_is_better": True,
        "scaling": "weak" if (args.shard == "config" and args.total_batch <= 0) else "strong",
        "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": shared_config(shape, len(q_all)),
        "run": {"step": f"one output bin = {P} batches of {B} walker configurations per GPU ({evals_per_step} evaluations over {world} GPU(s)), "
                        "bin folded" + (" and exchanged" if world > 1 else "") + " once per step",
                "batch_per_gpu": B, "batches_per_step": P, "evaluations_per_step": evals_per_step,
                "parallelism": (f"walker-configuration sharding x{world}, one NCCL reduce of the bin per step" if args.shard == "config"
                                else f"q-vector sharding x{world} ({nq} of {len(q_all)} q per GPU), one NCCL all-gather of the bin per step"),
                "l2": f"{use_slots} resident batches rotated ({footprint_mb:.0f} MB > 126 MB L2)",
                "rho_mode": args.rho_mode, "corr_mode": args.corr_mode,
                "collective": ("none (one GPU)" if world == 1 else
                               (("library NCCL, pipelined: pimcb_reduce_bins_begin / _end, the snapshot of bin k is reduced on the library's "
                                 "communication stream while bin k+1 is measured" if pipelined else
                                 "library NCCL (pimcb_reduce_bins / pimcb_gather_bins_q) on the compute stream") if lib_coll else "torch.distributed NCCL on the library's bin")),
                "host_binding": (f"rank bound to {len(numa_cpus)} GPU-local CPUs (NVML affinity)" if numa_cpus else "none")},
        "clocks": clocks, "e2e": e2e, "latency": latency, "gpu_launches": int(launches), "roofline": roofline, "pair_sums": pair, "ssf_direct": direct,
    }
    exchange_collect()
    beads_one = np.array(pinned[0].array[0])
    for pa in pinned:
        pa.free()
    ctx.close()
    if rank == 0 and world == 1 and not args.no_ab:
        roofline["upstream_gpu_ab"] = upstream_gpu_ab(shape, q_all, beads_one)
        ab = roofline["upstream_gpu_ab"]
        if ab and latency:
            ab["ours_single_configuration_evaluations_per_s"] = latency["evaluations_per_s"]
            ab["ours_device_resident_evaluations_per_s"] = value
            ab["speedup_single_call_vs_accumulate_pattern"] = latency["evaluations_per_s"] / ab["best_accumulate_evaluations_per_s"]
            ab["speedup_device_resident_vs_kernels"] = value / ab["best_kernels_evaluations_per_s"]
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        # the GPU work is over and its buffers are gone: the CPU arm gets every host core again; same sample and the same
        # warm-up + averaging as --impl reference
        os.sched_setaffinity(0, all_cpus)
        arm = CpuArm(shape, q_all, args.cpu_seconds)
        cpu_value, _ = time_cpu_arm(arm, 5, 1)
        line["cpu_baseline"] = {"value": cpu_value, "unit": UNIT, "cores": arm.cores, "kind": arm.kind,
                                "sample": arm.sample_text(), "one_core_value": 1.0 / arm.one_core(), "extrapolated": True,
                                "full_size_record": FULL_SIZE_RECORD,
                                "note": "the GPU/CPU ratio is ~10^4 algorithm (O(Nq M^2 N^2) pair loop vs the factorised O(Nq N M) form) "
                                        "and ~10^2 hardware; the roofline fraction, not this ratio, measures the kernels"}
    else:
        line["cpu_baseline"] = None
    if rank == 0:
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def pair_legs(args, ctx, shape, pinned, use_slots, B):
    """Per-slice pair-potential sums (Vint, gradVSquared on the odd slices of the gsf action, sepHist) and the virial slice
    sums of the same batches, against the measured gather roofline of their access pattern."""
    import math
    max_sep = math.sqrt(sum((L / 2.0) ** 2 for L in shape.side))
    Vt, dVt, d2Vt, drt = synth.aziz_table_numpy(max_sep, second=True)
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
    codec = ctx.table_codec_info()
    tile = os.environ.get("PIMCB_PAIR_TILE", "1") != "0"
    packed = tile and codec["vd_packed"]
    # table reads: every pair is visited once.  Packed tables (table_codec.h): ONE 32-byte sector per pair serves V and
    # dV/dr; verbatim tables: V on every slice + dV/dr on the odd slices of the gsf action
    reads = B * shape.M * pairs if packed else B * ((shape.M - shape.M // 2) * pairs + (shape.M // 2) * 2 * pairs)
    gather_peak = 289.0e9                              # profiles/r02a_gather_peak.txt: 32-byte sectors / s out of L2, footprint <= 53 MB
    pair = {"metric": "pair-potential action sums/s (Vint[M] + gradVSquared[odd slices] + sepHist[M][50] per configuration)",
            "kernel": ("pair_tile_kernel (32 x 32 tiles, every pair once, division-free exact index; two launches: force kernel on the "
                       "gradVSquared slices" + (" reading packed (V, dV/dr) sectors" if packed else "") + ", 64-register V-only kernel on the rest)")
                      if tile else "pair_sym_kernel (ring, every pair once)",
            "value": B / p_s, "unit": "configurations/s", "avg_launch_ms": p_s * 1e3, "launches_timed": pn,
            "configurations_per_launch": B, "ms_per_64_configurations": p_s * 1e3 * 64.0 / B,
            "table_entries": len(Vt), "table_mb_verbatim": 2 * 8 * len(Vt) / 1e6, "table_mb_read": (32 * codec["sectors"] / 1e6) if packed else 2 * 8 * len(Vt) / 1e6,
            "packed_tables": codec,
            "pairs_per_launch": B * shape.M * pairs, "pair_rate_g_per_s": B * shape.M * pairs / p_s / 1e9,
            "table_reads_per_launch": reads, "table_read_rate_g_per_s": reads / p_s / 1e9,
            "roofline": {"bound": "L2 sector rate of random table reads (one 32-byte sector per read)", "achieved": reads / p_s / 1e9,
                         "peak": gather_peak / 1e9, "unit": "G reads/s", "frac": reads / p_s / gather_peak,
                         "peak_source": "tools/micro/gather_peak.cu on this pool's B200 (profiles/r02a_gather_peak.txt): 289 G reads/s for "
                                        "footprints <= 53 MB, 140 G/s at 106 MB (the verbatim V + dV/dr tables of C2), 73 G/s from DRAM",
                         "note": "ncu (profiles/r02*_kernels.md, 64 configurations per launch): L2 hit rate 98 %, 0.12 GB DRAM reads per launch, issue "
                                 "slots 61 % busy -- bound by instruction issue / dependent FP64 latency, not by the reads",
                         "hbm_algorithmic": {"bytes": 32 * reads + B * 8 * shape.ndim * shape.N * shape.M,
                                             "achieved_gbs": (32 * reads + B * 8 * shape.ndim * shape.N * shape.M) / p_s / 1e9}}}
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
    vpacked = os.environ.get("PIMCB_VIRIAL_TILE", "1") != "0" and ctx.table_codec_info()["dd_packed"]
    vg = B * pairs * (shape.M if vpacked else shape.M + shape.M // 2)      # one sector per pair, or dV/dr everywhere + d2V/dr2 on the odd slices
    pair["virial_sums"] = {"metric": "virial slice sums/s (4 sums per slice; gsf action, window deltas from the host)",
                           "value": B / v_s, "unit": "configurations/s", "avg_launch_ms": v_s * 1e3, "launches_timed": vn,
                           "configurations_per_launch": B, "ms_per_64_configurations": v_s * 1e3 * 64.0 / B,
                           "kernel": "virial_tile_kernel" + (" (packed (dV/dr, d2V/dr2) sectors" if vpacked else " (verbatim tables")
                                     + "; two launches: T-matrix kernel on its slices, gV-only kernel on the rest)",
                           "table_reads_per_launch": vg, "table_read_rate_g_per_s": vg / v_s / 1e9,
                           "roofline": {"bound": "L2 sector rate of random 8-byte table reads", "achieved": vg / v_s / 1e9,
                                        "peak": gather_peak / 1e9, "unit": "G reads/s", "frac": vg / v_s / gather_peak}}
    return pair


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
