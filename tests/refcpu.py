"""ctypes loader of oracle/_ref/librefcpu<NDIM>d.so: the reference's OWN CPU function bodies (oracle/ref_cpu_shim.cpp).
Test infrastructure only."""
import ctypes as C
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_up = C.POINTER(C.c_uint)


def _p(a, t=_dp):
    return a.ctypes.data_as(t) if a is not None else None


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class RefCpu:
    def __init__(self, ndim):
        path = os.path.join(ROOT, "oracle", "_ref", f"librefcpu{ndim}d.so")
        if not os.path.exists(path):
            pytest.skip(f"{path} not built (needs the upstream tree: make -C oracle ref)")
        self.lib = C.CDLL(path)
        self.ndim = ndim
        assert self.lib.refcpu_ndim() == ndim
        L = self.lib
        L.refcpu_qvectors.argtypes = [C.c_char_p, C.c_char_p, _dp, _dp, C.c_int]
        L.refcpu_qvectors2.argtypes = [C.c_double, C.c_double, C.c_char_p, _dp, _dp, C.c_int, _ip, C.c_int]
        L.refcpu_ssf.argtypes = [_dp, _up, _dp, C.c_int, C.c_int, C.c_int, _dp, C.c_int, _dp]
        L.refcpu_isf.argtypes = [_dp, _dp, C.c_int, C.c_int, C.c_int, _dp, C.c_int, _dp]
        L.refcpu_ssf_cyl.argtypes = [_dp, _up, _dp, C.c_int, C.c_int, C.c_int, _dp, _ip, C.c_int, C.c_double, _dp]
        L.refcpu_write_array.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_char_p, C.c_int]
        L.refcpu_read_array.argtypes = [C.c_int, C.c_char_p, C.c_void_p, C.c_int, _ip, _ip]
        if ndim == 3:
            L.refcpu_action.argtypes = [C.c_int, _dp, _dp, C.c_int, C.c_int, C.c_int, _ip, C.c_double, C.c_double, C.c_double,
                                        C.c_int, _dp, _dp, C.c_int, _dp, _dp, _ip, _dp, _dp, _dp, _dp, _dp, _dp, _dp, _dp, C.c_double]

    def qvectors(self, qtype, text, side):
        side = _f64(side)
        n = self.lib.refcpu_qvectors(qtype.encode(), text.encode(), _p(side), None, 0)
        out = np.zeros((n, self.ndim))
        assert self.lib.refcpu_qvectors(qtype.encode(), text.encode(), _p(side), _p(out), n) == n
        return out

    def qvectors2(self, dq, qmax, geometry, side):
        side = _f64(side)
        out = np.zeros((20000, self.ndim))
        sizes = np.zeros(256, dtype=np.int32)
        ns = self.lib.refcpu_qvectors2(dq, qmax, geometry.encode(), _p(side), _p(out), len(out), _p(sizes, _ip), len(sizes))
        shells, k = [], 0
        for n in sizes[:ns]:
            shells.append(out[k:k + n].copy())
            k += n
        return shells

    def ssf(self, side, beads, N, q, periodic=None):
        side, beads, q = _f64(side), _f64(beads), _f64(q)
        per = np.ascontiguousarray(periodic if periodic is not None else np.ones(self.ndim), dtype=np.uint32)
        M, Next, _ = beads.shape
        out = np.zeros(len(q))
        assert self.lib.refcpu_ssf(_p(side), _p(per, _up), _p(beads), M, N, Next, _p(q), len(q), _p(out)) == 0
        return out

    def isf(self, side, beads, N, q):
        side, beads, q = _f64(side), _f64(beads), _f64(q)
        M, Next, _ = beads.shape
        out = np.zeros((len(q), M))
        assert self.lib.refcpu_isf(_p(beads), _p(side), M, N, Next, _p(q), len(q), _p(out)) == 0
        return out

    def ssf_cyl(self, side, beads, N, shells, maxR, periodic=None):
        side, beads = _f64(side), _f64(beads)
        per = np.ascontiguousarray(periodic if periodic is not None else np.ones(self.ndim), dtype=np.uint32)
        q = _f64(np.vstack(shells))
        sizes = np.array([len(s) for s in shells], dtype=np.int32)
        M, Next, _ = beads.shape
        out = np.zeros(len(shells))
        n1d = self.lib.refcpu_ssf_cyl(_p(side), _p(per, _up), _p(beads), M, N, Next, _p(q), _p(sizes, _ip), len(shells), maxR, _p(out))
        return out, n1d

    # the text of one state-file array by the upstream operator<< (kind 0 beads, 1 links, 2 worm.beads) and back
    _KIND = {0: (np.float64, None), 1: (np.int32, 2), 2: (np.uint32, 1)}

    def write_array(self, kind, arr) -> str:
        dt, _ = self._KIND[kind]
        a = np.ascontiguousarray(arr, dtype=dt)
        R, Cc = a.shape[0], a.shape[1]
        n = self.lib.refcpu_write_array(kind, a.ctypes.data, R, Cc, None, 0)
        buf = C.create_string_buffer(n)
        assert self.lib.refcpu_write_array(kind, a.ctypes.data, R, Cc, buf, n) == n
        return buf.value.decode()

    def read_array(self, kind, text):
        dt, w = self._KIND[kind]
        w = self.ndim if kind == 0 else w
        out = np.zeros(len(text) // 2 + 16, dtype=dt)
        R, Cc = C.c_int(0), C.c_int(0)
        rc = self.lib.refcpu_read_array(kind, text.encode(), out.ctypes.data, out.size, C.byref(R), C.byref(Cc))
        if rc != 0:
            raise ValueError(f"upstream operator>> failed ({rc})")
        shape = (R.value, Cc.value) if w == 1 else (R.value, Cc.value, w)
        return out[:R.value * Cc.value * w].reshape(shape)

    def action(self, side, beads, N, tau, lam, VF, GF, period, window=5, mu=0.0, next_links=None, year=1979, spring_k=0.0):
        """dict of everything LocalAction / EnergyEstimator / VirialEnergyEstimator return for one configuration."""
        side, beads = _f64(side), _f64(beads)
        M, Next, _ = beads.shape
        nl = np.ascontiguousarray(next_links, dtype=np.int32) if next_links is not None else None
        vf, gf = _f64(VF), _f64(GF)
        r = {"vint": np.zeros(M), "f2": np.zeros(M), "sephist": np.zeros((M, 50), dtype=np.int32), "vir": np.zeros((M, 4)),
             "dtau": np.zeros(M), "dlam": np.zeros(M), "d2tau": np.zeros(M), "vkc": np.zeros(M), "scalars": np.zeros(1),
             "energy": np.zeros(9), "virial": np.zeros(19)}
        rc = self.lib.refcpu_action(year, _p(side), _p(beads), M, N, Next, _p(nl, _ip), tau, lam, mu, window, _p(vf), _p(gf), period,
                                    _p(r["vint"]), _p(r["f2"]), _p(r["sephist"], _ip), _p(r["vir"]), _p(r["dtau"]), _p(r["dlam"]),
                                    _p(r["d2tau"]), _p(r["vkc"]), _p(r["scalars"]), _p(r["energy"]), _p(r["virial"]), spring_k)
        assert rc == 0
        r["potentialAction"] = float(r.pop("scalars")[0])
        return r
