"""ctypes binding of libpimc_b200.so (include/pimc_b200.h).

Thin by design: one Python method per C entry point, numpy arrays in and out, errors raised as
`PimcbError` carrying pimcb_last_error().  There is no CPU fallback: if the library or a CUDA device is
missing the constructor raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import build as _build

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_up = C.POINTER(C.c_uint)
NPCFSEP = 50

# every symbol include/pimc_b200.h declares (checked by tests/test_abi.py)
SYMBOLS = [
    "pimcb_create", "pimcb_destroy", "pimcb_last_error", "pimcb_version", "pimcb_set_box", "pimcb_set_qvecs",
    "pimcb_num_commensurate", "pimcb_set_rho_mode", "pimcb_set_corr_mode", "pimcb_stage_beads", "pimcb_stage_batch", "pimcb_stage_batch_async", "pimcb_stage_wait", "pimcb_num_slots",
    "pimcb_stage_batch_slot", "pimcb_select_slot", "pimcb_host_alloc", "pimcb_host_free", "pimcb_host_register",
    "pimcb_host_unregister", "pimcb_ssf", "pimcb_isf", "pimcb_ssf_isf", "pimcb_ssf_isf_beads", "pimcb_measure", "pimcb_reset_bins",
    "pimcb_read_bins", "pimcb_bins_device_ptr", "pimcb_sync", "pimcb_stream", "pimcb_set_pair_table",
    "pimcb_pair_sums", "pimcb_measure_fp64_peak", "pimcb_set_profiling", "pimcb_set_profiling_stride", "pimcb_kernel_times",
    "pimcb_launch_count", "pimcb_rho_plan_info", "pimcb_elastic", "pimcb_ssf_cyl", "pimcb_set_pair_table_d2",
    "pimcb_virial_sums", "pimcb_comm_unique_id", "pimcb_comm_init", "pimcb_comm_destroy", "pimcb_reduce_bins",
    "pimcb_reduce_bins_begin", "pimcb_reduce_bins_end",
    "pimcb_gather_bins_q", "pimcb_set_external_gradient", "pimcb_set_external_laplacian", "pimcb_init_bins", "pimcb_measure_h2d_peak", "pimcb_table_codec_info",
]


class PimcbError(RuntimeError):
    pass


def _prefer_bundled_nccl() -> None:
    """The library dlopens NCCL by soname.  In a Python process that also imports torch, the copy bundled with torch's
    wheels (nvidia/nccl/lib) must be the one in the process: a system libnccl.so.2 loaded first would satisfy torch's own
    DT_NEEDED by soname and miss symbols torch was built against.  Point PIMCB_NCCL_LIB at the bundled copy (found
    without importing torch) unless the caller chose one."""
    if os.environ.get("PIMCB_NCCL_LIB"):
        return
    import importlib.util
    try:
        spec = importlib.util.find_spec("nvidia.nccl")
    except (ImportError, ValueError):
        spec = None
    for root in (list(spec.submodule_search_locations) if spec and spec.submodule_search_locations else []):
        cand = os.path.join(root, "lib", "libnccl.so.2")
        if os.path.exists(cand):
            os.environ["PIMCB_NCCL_LIB"] = cand
            return


_lib = None


def load_library(path: str | None = None) -> C.CDLL:
    """dlopen the in-tree library (never builds implicitly on import; build via __graft_entry__.build())."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    path = path or _build.LIB
    if not os.path.exists(path):
        raise PimcbError(f"{path} is missing: run `python -m pimc_b200.build` (nvcc, sm_100a) first")
    lib = C.CDLL(path)
    vp = C.c_void_p
    lib.pimcb_last_error.restype = C.c_char_p
    lib.pimcb_launch_count.restype = C.c_long
    lib.pimcb_launch_count.argtypes = [vp]
    lib.pimcb_create.argtypes = [C.POINTER(vp), C.c_int, C.c_int]
    lib.pimcb_destroy.argtypes = [vp]
    lib.pimcb_set_box.argtypes = [vp, _dp, _up]
    lib.pimcb_set_qvecs.argtypes = [vp, _dp, C.c_int]
    lib.pimcb_num_commensurate.argtypes = [vp]
    lib.pimcb_set_rho_mode.argtypes = [vp, C.c_int]
    lib.pimcb_set_corr_mode.argtypes = [vp, C.c_int]
    lib.pimcb_stage_beads.argtypes = [vp, _dp, C.c_int, C.c_int, C.c_int]
    lib.pimcb_stage_batch.argtypes = [vp, _dp, C.c_int, C.c_int, C.c_int, C.c_int]
    lib.pimcb_stage_batch_async.argtypes = [vp, _dp, C.c_int, C.c_int, C.c_int, C.c_int]
    lib.pimcb_stage_wait.argtypes = [vp]
    lib.pimcb_num_slots.argtypes = [vp]
    lib.pimcb_stage_batch_slot.argtypes = [vp, C.c_int, _dp, C.c_int, C.c_int, C.c_int, C.c_int]
    lib.pimcb_select_slot.argtypes = [vp, C.c_int]
    lib.pimcb_host_alloc.argtypes = [C.POINTER(vp), C.c_size_t]
    lib.pimcb_host_free.argtypes = [vp]
    lib.pimcb_host_register.argtypes = [vp, C.c_size_t]
    lib.pimcb_host_unregister.argtypes = [vp]
    lib.pimcb_ssf.argtypes = [vp, _dp]
    lib.pimcb_isf.argtypes = [vp, _dp]
    lib.pimcb_ssf_isf.argtypes = [vp, _dp, _dp]
    lib.pimcb_ssf_isf_beads.argtypes = [vp, _dp, C.c_int, C.c_int, C.c_int, _dp, _dp]
    lib.pimcb_measure.argtypes = [vp]
    lib.pimcb_reset_bins.argtypes = [vp]
    lib.pimcb_read_bins.argtypes = [vp, _dp, _dp, C.POINTER(C.c_long)]
    lib.pimcb_bins_device_ptr.argtypes = [vp, C.POINTER(vp), C.POINTER(C.c_size_t)]
    lib.pimcb_sync.argtypes = [vp]
    lib.pimcb_stream.argtypes = [vp, C.POINTER(vp)]
    lib.pimcb_set_pair_table.argtypes = [vp, _dp, _dp, C.c_int, C.c_double, _dp, _dp]
    lib.pimcb_pair_sums.argtypes = [vp, _dp, _dp, _ip, C.c_double, C.c_int]
    lib.pimcb_measure_fp64_peak.argtypes = [vp, _dp, C.c_double]
    lib.pimcb_set_profiling.argtypes = [vp, C.c_int]
    lib.pimcb_set_profiling_stride.argtypes = [vp, C.c_int]
    lib.pimcb_kernel_times.argtypes = [vp, _dp, C.POINTER(C.c_long), C.c_int]
    lib.pimcb_rho_plan_info.argtypes = [vp, _ip]
    lib.pimcb_elastic.argtypes = [vp, _dp]
    lib.pimcb_ssf_cyl.argtypes = [vp, C.c_double, _dp, _ip]
    lib.pimcb_set_pair_table_d2.argtypes = [vp, _dp, C.c_int, _dp]
    lib.pimcb_virial_sums.argtypes = [vp, _dp, C.c_int, _dp]
    lib.pimcb_set_external_gradient.argtypes = [vp, _dp]
    lib.pimcb_set_external_laplacian.argtypes = [vp, _dp]
    lib.pimcb_init_bins.argtypes = [vp, C.c_int]
    lib.pimcb_table_codec_info.argtypes = [vp, C.POINTER(C.c_long)]
    lib.pimcb_measure_h2d_peak.argtypes = [vp, C.c_void_p, C.c_size_t, C.c_int, _dp]
    lib.pimcb_comm_unique_id.argtypes = [C.c_char_p]
    lib.pimcb_comm_init.argtypes = [vp, C.c_int, C.c_int, C.c_char_p]
    lib.pimcb_comm_destroy.argtypes = [vp]
    lib.pimcb_reduce_bins.argtypes = [vp, C.c_int, C.POINTER(C.c_long)]
    lib.pimcb_reduce_bins_begin.argtypes = [vp, C.c_int]
    lib.pimcb_reduce_bins_end.argtypes = [vp, _dp, _dp, C.POINTER(C.c_long)]
    lib.pimcb_gather_bins_q.argtypes = [vp, _ip, _dp, _dp]
    if path == _build.LIB:
        _lib = lib
    return lib


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _ptr(a):
    return a.ctypes.data_as(_dp) if a is not None else None


class PinnedArray:
    """A numpy view over pimcb_host_alloc'd (page-locked) memory."""

    def __init__(self, shape, lib=None):
        self._lib = lib or load_library()
        n = int(np.prod(shape))
        self._p = C.c_void_p()
        rc = self._lib.pimcb_host_alloc(C.byref(self._p), n * 8)
        if rc:
            raise PimcbError(self._lib.pimcb_last_error().decode())
        buf = (C.c_double * n).from_address(self._p.value)
        self.array = np.frombuffer(buf, dtype=np.float64).reshape(shape)

    def free(self):
        if self._p:
            self.array = None
            self._lib.pimcb_host_free(self._p)
            self._p = None


class Context:
    """One pimcb_ctx: one device, one spatial dimension, one box + q-set."""

    KERNELS = ("rho", "corr", "ssf_direct", "bins", "pair", "transpose", "variant", "virial")

    def __init__(self, device: int = 0, ndim: int = 3):
        self.lib = load_library()
        self.ndim = ndim
        self._h = C.c_void_p()
        self._chk(self.lib.pimcb_create(C.byref(self._h), device, ndim))
        self.nq = 0
        self.shape = None   # (B, M, N) of the current slot
        self._slot_shapes = {}

    # -- plumbing ---------------------------------------------------------------------------
    def _chk(self, rc):
        if rc != 0:
            raise PimcbError(f"pimcb error {rc}: {self.lib.pimcb_last_error().decode()}")

    def close(self):
        if getattr(self, "_h", None):
            self.lib.pimcb_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- configuration -------------------------------------------------------------------------
    def set_box(self, side, periodic=None):
        side = _f64(side)
        assert side.shape == (self.ndim,)
        per = None
        if periodic is not None:
            per = np.ascontiguousarray(periodic, dtype=np.uint32)
        self._chk(self.lib.pimcb_set_box(self._h, _ptr(side), per.ctypes.data_as(_up) if per is not None else None))

    def set_qvecs(self, q):
        q = _f64(q)
        assert q.ndim == 2 and q.shape[1] == self.ndim
        self._chk(self.lib.pimcb_set_qvecs(self._h, _ptr(q), len(q)))
        self.nq = len(q)

    def num_commensurate(self) -> int:
        return self.lib.pimcb_num_commensurate(self._h)

    def set_rho_mode(self, mode: int):
        self._chk(self.lib.pimcb_set_rho_mode(self._h, mode))

    def set_corr_mode(self, mode: int):
        self._chk(self.lib.pimcb_set_corr_mode(self._h, mode))

    # -- staging -----------------------------------------------------------------------------------
    @staticmethod
    def _batch(beads, N):
        beads = np.asarray(beads)
        if beads.dtype != np.float64 or not beads.flags.c_contiguous:
            beads = _f64(beads)
        if beads.ndim == 3:
            beads = beads[None]
        B, M, Next, nd = beads.shape
        return beads, B, M, Next, nd

    def stage(self, beads, N: int, slot: int | None = None):
        """beads: [M][N_ext][ndim] or [B][M][N_ext][ndim] float64 (reference AoS layout)."""
        beads, B, M, Next, nd = self._batch(beads, N)
        assert nd == self.ndim
        if slot is None:
            self._chk(self.lib.pimcb_stage_batch(self._h, _ptr(beads), B, M, N, Next))
            self.shape = (B, M, N)
        else:
            self._chk(self.lib.pimcb_stage_batch_slot(self._h, slot, _ptr(beads), B, M, N, Next))
            self._slot_shapes[slot] = (B, M, N)
        return self

    def stage_async(self, beads, N: int):
        """Page-locked source: enqueue the DMA and return; keep `beads` untouched until stage_wait() / results are read."""
        beads, B, M, Next, nd = self._batch(beads, N)
        assert nd == self.ndim
        self._chk(self.lib.pimcb_stage_batch_async(self._h, _ptr(beads), B, M, N, Next))
        self.shape = (B, M, N)
        return self

    def stage_wait(self):
        self._chk(self.lib.pimcb_stage_wait(self._h))

    def select_slot(self, slot: int):
        self._chk(self.lib.pimcb_select_slot(self._h, slot))
        self.shape = self._slot_shapes.get(slot, self.shape)

    def num_slots(self) -> int:
        return self.lib.pimcb_num_slots(self._h)

    def ssf_isf_beads(self, beads, N: int):
        """One configuration [M][N_ext][ndim]: stage + S(q) + F(q,tau) + read-back with a single synchronisation."""
        beads, B, M, Next, nd = self._batch(beads, N)
        assert B == 1 and nd == self.ndim
        ssf = np.zeros((1, self.nq))
        isf = np.zeros((1, self.nq, M))
        self._chk(self.lib.pimcb_ssf_isf_beads(self._h, _ptr(beads), M, N, Next, _ptr(ssf), _ptr(isf)))
        self.shape = (1, M, N)
        return ssf, isf

    # -- estimators ------------------------------------------------------------------------------------
    def ssf_isf(self, want_ssf=True, want_isf=True):
        B, M, _ = self.shape
        ssf = np.zeros((B, self.nq)) if want_ssf else None
        isf = np.zeros((B, self.nq, M)) if want_isf else None
        self._chk(self.lib.pimcb_ssf_isf(self._h, _ptr(ssf), _ptr(isf)))
        return ssf, isf

    def ssf(self):
        B, _, _ = self.shape
        out = np.zeros((B, self.nq))
        self._chk(self.lib.pimcb_ssf(self._h, _ptr(out)))
        return out

    def isf(self):
        B, M, _ = self.shape
        out = np.zeros((B, self.nq, M))
        self._chk(self.lib.pimcb_isf(self._h, _ptr(out)))
        return out

    def measure(self):
        self._chk(self.lib.pimcb_measure(self._h))

    def reset_bins(self):
        self._chk(self.lib.pimcb_reset_bins(self._h))

    def read_bins(self):
        M = self.shape[1] if getattr(self, "shape", None) else self._bins_M     # nothing staged yet: the init_bins layout
        ssf = np.zeros(self.nq)
        isf = np.zeros((self.nq, M))
        n = C.c_long(0)
        self._chk(self.lib.pimcb_read_bins(self._h, _ptr(ssf), _ptr(isf), C.byref(n)))
        return ssf, isf, n.value

    def bins_device_ptr(self):
        p, n = C.c_void_p(), C.c_size_t(0)
        self._chk(self.lib.pimcb_bins_device_ptr(self._h, C.byref(p), C.byref(n)))
        return p.value, n.value

    def sync(self):
        self._chk(self.lib.pimcb_sync(self._h))

    def stream(self) -> int:
        p = C.c_void_p()
        self._chk(self.lib.pimcb_stream(self._h, C.byref(p)))
        return p.value or 0

    # -- pair potential --------------------------------------------------------------------------------
    def set_pair_table(self, V, dVdr, dr, extV=(0.0, 0.0), extdVdr=(0.0, 0.0)):
        V = _f64(V)
        dV = _f64(dVdr) if dVdr is not None else None
        e0, e1 = _f64(extV), _f64(extdVdr)
        self._chk(self.lib.pimcb_set_pair_table(self._h, _ptr(V), _ptr(dV), len(V), dr, _ptr(e0), _ptr(e1)))

    def set_external_gradient(self, gext):
        """gext: gradient of the external potential per bead, shaped like the staged beads (or None to clear)."""
        g = _f64(gext) if gext is not None else None
        self._chk(self.lib.pimcb_set_external_gradient(self._h, _ptr(g)))

    def h2d_peak_gbs(self, pinned_array: np.ndarray, reps: int = 8) -> float:
        """Bare cudaMemcpyAsync rate of a page-locked array on this context's copy stream (GB/s)."""
        g = C.c_double(0.0)
        self._chk(self.lib.pimcb_measure_h2d_peak(self._h, pinned_array.ctypes.data_as(C.c_void_p), pinned_array.nbytes, int(reps),
                                                  C.byref(g)))
        return g.value

    def table_codec_info(self) -> dict:
        v = (C.c_long * 5)()
        self._chk(self.lib.pimcb_table_codec_info(self._h, v))
        return {"vd_packed": bool(v[0]), "vd_raw_sectors": v[1], "dd_packed": bool(v[2]), "dd_raw_sectors": v[3], "sectors": v[4]}

    def init_bins(self, M: int):
        """Zeroed bin for M slices before any measurement (lets an idle rank join the bin collective)."""
        self._chk(self.lib.pimcb_init_bins(self._h, int(M)))
        self._bins_M = int(M)

    def set_external_laplacian(self, g2ext):
        """g2ext: Laplacian of the external potential per bead, [B][M][N_ext] (or [M][N_ext]; None to clear)."""
        g = _f64(g2ext) if g2ext is not None else None
        self._chk(self.lib.pimcb_set_external_laplacian(self._h, _ptr(g)))

    def pair_sums(self, dSep=None, want_f2=True, want_hist=True, f2_parity=-1):
        B, M, _ = self.shape
        vint = np.zeros((B, M))
        f2 = np.zeros((B, M)) if want_f2 else None
        hist = np.zeros((B, M, NPCFSEP), dtype=np.int32) if want_hist else None
        self._chk(self.lib.pimcb_pair_sums(self._h, _ptr(vint), _ptr(f2),
                                           hist.ctypes.data_as(_ip) if hist is not None else None,
                                           float(dSep) if dSep else 0.0, f2_parity))
        return vint, f2, hist

    # -- scattering variants / virial ---------------------------------------------------------------------
    def elastic(self):
        """[B][nq]: the upstream elastic-scattering estimator's per-measurement increment."""
        B, _, _ = self.shape
        out = np.zeros((B, self.nq))
        self._chk(self.lib.pimcb_elastic(self._h, _ptr(out)))
        return out

    def ssf_cyl(self, maxR: float):
        """([B][nq] raw cylinder S(q) sums, [B] beads of slice 0 inside the radius)."""
        B, _, _ = self.shape
        out = np.zeros((B, self.nq))
        n_in = np.zeros(B, dtype=np.int32)
        self._chk(self.lib.pimcb_ssf_cyl(self._h, float(maxR), _ptr(out), n_in.ctypes.data_as(_ip)))
        return out, n_in

    def set_pair_table_d2(self, d2V, ext=(0.0, 0.0)):
        d2V, e = _f64(d2V), _f64(ext)
        self._chk(self.lib.pimcb_set_pair_table_d2(self._h, _ptr(d2V), len(d2V), _ptr(e)))

    def virial_sums(self, delta=None, t2_parity=-1):
        """[B][M][4] = {sum gV.r, sum (T gV).r, sum gV.delta, sum (T gV).delta}; delta in the staged beads' AoS shape."""
        B, M, _ = self.shape
        d = _f64(delta) if delta is not None else None
        out = np.zeros((B, M, 4))
        self._chk(self.lib.pimcb_virial_sums(self._h, _ptr(d), t2_parity, _ptr(out)))
        return out

    # -- multi-GPU exchange step (NCCL inside the library) ---------------------------------------------------
    @staticmethod
    def comm_unique_id() -> bytes:
        _prefer_bundled_nccl()
        lib = load_library()
        buf = C.create_string_buffer(128)
        rc = lib.pimcb_comm_unique_id(buf)
        if rc:
            raise PimcbError(f"pimcb error {rc}: {lib.pimcb_last_error().decode()}")
        return buf.raw

    def comm_init(self, nranks: int, rank: int, unique_id: bytes):
        assert len(unique_id) == 128
        _prefer_bundled_nccl()
        self._chk(self.lib.pimcb_comm_init(self._h, nranks, rank, unique_id))

    def comm_destroy(self):
        self._chk(self.lib.pimcb_comm_destroy(self._h))

    def reduce_bins(self, root: int = 0, want_total: bool = True) -> int:
        """want_total=False: everything is enqueued, nothing waited for; read_bins() returns the global count later."""
        if not want_total:
            self._chk(self.lib.pimcb_reduce_bins(self._h, root, None))
            return 0
        n = C.c_long(0)
        self._chk(self.lib.pimcb_reduce_bins(self._h, root, C.byref(n)))
        return n.value

    def reduce_bins_begin(self, root: int = 0):
        """Snapshot the bin and start its reduce on the library's communication stream; reset_bins() and the next bin's
        measurements may follow at once."""
        M = self.shape[1] if getattr(self, "shape", None) else self._bins_M     # nothing staged yet: the init_bins layout
        self._xchg = (self.nq, M)
        self._chk(self.lib.pimcb_reduce_bins_begin(self._h, root))
        return self

    def reduce_bins_end(self):
        """Wait for the exchange started by reduce_bins_begin; (ssf[nq], isf[nq][M], configurations) -- zeros off the root."""
        nq, M = getattr(self, "_xchg", None) or (self.nq, self.shape[1] if getattr(self, "shape", None) else self._bins_M)
        ssf, isf, n = np.zeros(nq), np.zeros((nq, M)), C.c_long(0)
        self._chk(self.lib.pimcb_reduce_bins_end(self._h, _ptr(ssf), _ptr(isf), C.byref(n)))
        return ssf, isf, n.value

    def gather_bins_q(self, nq_per_rank):
        _, M, _ = self.shape
        sizes = np.ascontiguousarray(nq_per_rank, dtype=np.int32)
        tot = int(sizes.sum())
        ssf, isf = np.zeros(tot), np.zeros((tot, M))
        self._chk(self.lib.pimcb_gather_bins_q(self._h, sizes.ctypes.data_as(_ip), _ptr(ssf), _ptr(isf)))
        return ssf, isf

    # -- measurement helpers ----------------------------------------------------------------------------
    def fp64_peak_tflops(self, seconds=0.5) -> float:
        v = C.c_double(0.0)
        self._chk(self.lib.pimcb_measure_fp64_peak(self._h, C.byref(v), seconds))
        return v.value

    KERNEL_IDS = {"rho": 0, "corr": 1, "direct": 2, "bins": 3, "pair": 4, "transpose": 5, "variant": 6, "virial": 7}

    def set_profiling(self, on):
        """False / True (every kernel) or an iterable of kernel names to time."""
        if isinstance(on, (bool, int)):
            mode = int(bool(on))
        else:
            mode = sum(1 << (self.KERNEL_IDS[k] + 1) for k in on)
        self._chk(self.lib.pimcb_set_profiling(self._h, mode))

    def set_profiling_stride(self, stride: int):
        self._chk(self.lib.pimcb_set_profiling_stride(self._h, int(stride)))

    def kernel_times(self, reset: bool = True) -> dict:
        """{kernel: (total_ms, launches)} since the last reset (needs set_profiling(True))."""
        ms = (C.c_double * 8)()
        cnt = (C.c_long * 8)()
        self._chk(self.lib.pimcb_kernel_times(self._h, ms, cnt, int(reset)))
        return {k: (ms[i], cnt[i]) for i, k in enumerate(self.KERNELS)}

    def rho_plan_info(self) -> dict:
        v = (C.c_int * 12)()
        self._chk(self.lib.pimcb_rho_plan_info(self._h, v))
        keys = ("path", "groups", "L_rows", "R_cols", "M_tiles", "N_tiles", "nmax_x", "nmax_y", "nmax_z", "commensurate",
                "non_commensurate", "nq")
        d = dict(zip(keys, list(v)))
        d["path_name"] = {0: "rho_generic_kernel", 1: "rho_lattice_mma_kernel", 2: "rho_lattice_kernel"}.get(d["path"], "none")
        return d

    def launch_count(self) -> int:
        return self.lib.pimcb_launch_count(self._h)
