"""ctypes loader for the CPU oracle (oracle/pimc_oracle.cpp).

TEST INFRASTRUCTURE ONLY.  Importable from tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py; the product package pimc_b200 never
imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_up = C.POINTER(C.c_uint)

NPCFSEP = 50


def build(force: bool = False) -> None:
    """Compile liboracle.so / liboracle_fast.so with oracle/Makefile (g++ only)."""
    need = force or not all(os.path.exists(os.path.join(_HERE, n)) for n in ("liboracle.so", "liboracle_fast.so"))
    if need:
        subprocess.run(["make", "-C", _HERE] + (["-B"] if force else []), check=True,
                       stdout=subprocess.DEVNULL)


def _dptr(a):
    return a.ctypes.data_as(_dp) if a is not None else None


def _iptr(a):
    return a.ctypes.data_as(_ip) if a is not None else None


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class Oracle:
    """Thin wrapper over one build of the oracle library (`fast=True` = reference optimisation flags)."""

    def __init__(self, fast: bool = False):
        build()
        self.lib = C.CDLL(os.path.join(_HERE, "liboracle_fast.so" if fast else "liboracle.so"))
        L = self.lib
        L.orc_max_sep.restype = C.c_double
        L.orc_max_sep.argtypes = [C.c_int, _dp, _up]
        L.orc_aziz_rm.restype = C.c_double
        L.orc_aziz_rm.argtypes = [C.c_int]
        L.orc_potential_action.restype = C.c_double
        L.orc_potential_action.argtypes = [_dp, _dp, C.c_int, _dp, _dp, C.c_double, C.c_double]
        L.orc_deriv_potential_action_tau.restype = C.c_double
        L.orc_deriv_potential_action_tau.argtypes = [C.c_double, C.c_double, C.c_int, _dp, _dp, C.c_double, C.c_double]
        L.orc_deriv_potential_action_lambda.restype = C.c_double
        L.orc_deriv_potential_action_lambda.argtypes = [C.c_double, C.c_int, _dp, C.c_double]
        L.orc_aziz_tail.restype = C.c_double
        L.orc_aziz_tail.argtypes = [C.c_int, C.c_double]
        L.orc_energy.argtypes = [C.c_int, _dp, _up, _dp, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), _dp, _dp, _dp, _dp,
                                 C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, _dp]
        L.orc_time_slices.restype = C.c_int
        L.orc_time_slices.argtypes = [C.c_double, C.c_double, C.c_int, _dp]
        L.orc_qvectors.argtypes = [C.c_int, C.c_char_p, C.c_char_p, _dp, _dp, C.c_int]
        L.orc_ssf.argtypes = [C.c_int, _dp, _up, _dp, C.c_int, C.c_int, C.c_int, _dp, C.c_int, _dp]
        L.orc_ssf_mt.argtypes = L.orc_ssf.argtypes + [C.c_int]
        L.orc_isf.argtypes = [C.c_int, _dp, C.c_int, C.c_int, C.c_int, _dp, C.c_int, _dp, C.c_int]
        L.orc_isf_range.argtypes = [C.c_int, _dp, C.c_int, C.c_int, C.c_int, _dp, C.c_int, _dp, C.c_long, C.c_long, C.c_int]
        L.orc_isf_factorised.argtypes = [C.c_int, _dp, C.c_int, C.c_int, C.c_int, _dp, C.c_int, _dp]
        L.orc_aziz_values.argtypes = [C.c_int, C.c_int, _dp, _dp, C.c_int]
        L.orc_aziz_table.argtypes = [C.c_int, C.c_double, _dp, _dp, _dp, C.c_int, _dp]
        L.orc_table_V.argtypes = [C.c_int, _dp, C.c_int, C.c_double, _dp, _dp, _dp, C.c_int]
        L.orc_pair_sums.argtypes = [C.c_int, _dp, _up, _dp, C.c_int, C.c_int, C.c_int, _dp, _dp, C.c_int,
                                    C.c_double, _dp, _dp, C.c_double, _dp, _dp, _ip, C.c_int]
        L.orc_grad_v_squared_ext.argtypes = [C.c_int, _dp, _up, _dp, C.c_int, C.c_int, C.c_int, _dp, C.c_int, C.c_double, _dp, _dp, _dp]
        L.orc_put_in_bc.argtypes = [C.c_int, _dp, _up, _dp, C.c_int]
        L.orc_format_row.argtypes = [_dp, _dp, C.c_int, C.c_uint, C.c_char_p, C.c_int]
        L.orc_dvec_to_string.argtypes = [C.c_int, _dp, C.c_char_p, C.c_int]
        L.orc_qvectors2.argtypes = [C.c_int, C.c_double, C.c_double, C.c_char_p, _dp, C.c_int, _ip, C.c_int, _ip]
        L.orc_ssf_cyl.argtypes = [C.c_int, _dp, _up, _dp, C.c_int, C.c_int, C.c_int, _dp, C.c_int, C.c_double, _dp, _ip]
        L.orc_elastic.argtypes = [C.c_int, _dp, C.c_int, C.c_int, C.c_int, _dp, C.c_int, _dp, C.c_int]
        L.orc_virial_delta.argtypes = [C.c_int, _dp, _up, _dp, C.c_int, C.c_int, C.c_int, _ip, C.c_int, _dp]
        L.orc_virial_sums.argtypes = [C.c_int, _dp, _up, _dp, C.c_int, C.c_int, C.c_int, _ip, C.c_int, _dp, _dp, C.c_int,
                                      C.c_double, _dp, _dp, C.c_int, _dp, C.c_int]
        L.orc_virial_sums_ext.argtypes = [C.c_int, _dp, _up, _dp, C.c_int, C.c_int, C.c_int, _ip, C.c_int, _dp, _dp, C.c_int,
                                          C.c_double, _dp, _dp, C.c_int, _dp, C.c_int, _dp, _dp]
        L.orc_virial_energy.argtypes = [C.c_int, _dp, _up, _dp, C.c_int, C.c_int, C.c_int, _ip, C.c_int, _dp, _dp, _dp, _dp, _dp,
                                        C.c_double, C.c_double, C.c_double, C.c_double, C.c_int, _dp]

    # -- geometry ------------------------------------------------------------------
    @staticmethod
    def _box(side, periodic):
        side = _f64(side)
        per = np.ascontiguousarray(periodic if periodic is not None else np.ones(len(side)), dtype=np.uint32)
        return side, per

    def max_sep(self, side, periodic=None) -> float:
        side, per = self._box(side, periodic)
        return self.lib.orc_max_sep(len(side), _dptr(side), per.ctypes.data_as(_up))

    def put_in_bc(self, side, r, periodic=None):
        side, per = self._box(side, periodic)
        r = _f64(r).copy()
        rc = self.lib.orc_put_in_bc(len(side), _dptr(side), per.ctypes.data_as(_up), _dptr(r), r.size // len(side))
        assert rc == 0
        return r

    def time_slices(self, T, tau=0.0, P=0):
        t = C.c_double(0.0)
        M = self.lib.orc_time_slices(T, tau, P, C.byref(t))
        return M, t.value

    # -- wave-vectors --------------------------------------------------------------
    def qvectors(self, qtype: str, text: str, side) -> np.ndarray:
        side = _f64(side)
        nd = len(side)
        n = self.lib.orc_qvectors(nd, qtype.encode(), text.encode(), _dptr(side), None, 0)
        if n < 0:
            raise ValueError(f"orc_qvectors failed with {n}")
        out = np.zeros((n, nd))
        n2 = self.lib.orc_qvectors(nd, qtype.encode(), text.encode(), _dptr(side), _dptr(out), n)
        assert n2 == n
        return out

    # -- estimators ----------------------------------------------------------------
    @staticmethod
    def _beads(beads):
        beads = _f64(beads)
        M, Next, nd = beads.shape
        return beads, M, Next, nd

    def ssf(self, side, beads, N, q, periodic=None, nthreads=1) -> np.ndarray:
        """sf(q)/N as added to the estimator by one accumulate() call."""
        beads, M, Next, nd = self._beads(beads)
        side, per = self._box(side, periodic)
        q = _f64(q)
        out = np.zeros(len(q))
        rc = self.lib.orc_ssf_mt(nd, _dptr(side), per.ctypes.data_as(_up), _dptr(beads), M, N, Next, _dptr(q),
                                 len(q), _dptr(out), nthreads)
        assert rc == 0
        return out

    def isf(self, beads, N, q, nthreads=1) -> np.ndarray:
        """isf[q, tau]/N by the reference's direct O(Nq M^2 N^2) loop."""
        beads, M, Next, nd = self._beads(beads)
        q = _f64(q)
        out = np.zeros((len(q), M))
        rc = self.lib.orc_isf(nd, _dptr(beads), M, N, Next, _dptr(q), len(q), _dptr(out), nthreads)
        assert rc == 0
        return out

    def isf_range(self, beads, N, q, e0, e1, nthreads=1) -> np.ndarray:
        """Flattened elements [e0,e1) of isf[q*M+tau]/N (others left 0); bounded samples for CPU timing."""
        beads, M, Next, nd = self._beads(beads)
        q = _f64(q)
        out = np.zeros(len(q) * M)
        rc = self.lib.orc_isf_range(nd, _dptr(beads), M, N, Next, _dptr(q), len(q), _dptr(out), e0, e1, nthreads)
        assert rc == 0
        return out

    def isf_factorised(self, beads, N, q) -> np.ndarray:
        beads, M, Next, nd = self._beads(beads)
        q = _f64(q)
        out = np.zeros((len(q), M))
        rc = self.lib.orc_isf_factorised(nd, _dptr(beads), M, N, Next, _dptr(q), len(q), _dptr(out))
        assert rc == 0
        return out

    # -- pair potential --------------------------------------------------------------
    def aziz_rm(self, year=1979) -> float:
        return self.lib.orc_aziz_rm(year)

    def aziz_values(self, r, which=0, year=1979) -> np.ndarray:
        r = _f64(r)
        out = np.zeros_like(r)
        self.lib.orc_aziz_values(year, which, _dptr(r), _dptr(out), r.size)
        return out

    def aziz_table(self, max_sep, year=1979, second=False):
        """(V, dVdr[, d2Vdr2], dr) lookup tables exactly as TabulatedPotential::initLookupTable builds them."""
        dr = C.c_double(0.0)
        n = self.lib.orc_aziz_table(year, max_sep, None, None, None, 0, C.byref(dr))
        V = np.zeros(n)
        dV = np.zeros(n)
        d2V = np.zeros(n) if second else None
        n2 = self.lib.orc_aziz_table(year, max_sep, _dptr(V), _dptr(dV), _dptr(d2V), n, C.byref(dr))
        assert n2 == n
        return (V, dV, d2V, dr.value) if second else (V, dV, dr.value)

    def table_V(self, table, dr, sep, ext=(0.0, 0.0)) -> np.ndarray:
        sep = _f64(sep)
        nd = sep.shape[-1]
        table = _f64(table)
        ext = _f64(ext)
        out = np.zeros(sep.shape[0])
        rc = self.lib.orc_table_V(nd, _dptr(table), len(table), dr, _dptr(ext), _dptr(sep), _dptr(out), len(out))
        assert rc == 0
        return out

    def pair_sums(self, side, beads, N, V, dVdr, dr, dSep, periodic=None, want_f2=True, want_hist=True,
                  nthreads=1):
        """Per-slice (Vint[M], gradVSquared[M] | None, sepHist[M,50] | None)."""
        beads, M, Next, nd = self._beads(beads)
        side, per = self._box(side, periodic)
        V = _f64(V)
        dVdr = _f64(dVdr) if dVdr is not None else None
        ext = np.zeros(2)
        vint = np.zeros(M)
        f2 = np.zeros(M) if (want_f2 and dVdr is not None) else None
        hist = np.zeros((M, NPCFSEP), dtype=np.int32) if want_hist else None
        rc = self.lib.orc_pair_sums(nd, _dptr(side), per.ctypes.data_as(_up), _dptr(beads), M, N, Next, _dptr(V),
                                    _dptr(dVdr), len(V), dr, _dptr(ext), _dptr(ext), dSep, _dptr(vint), _dptr(f2),
                                    _iptr(hist), nthreads)
        assert rc == 0
        return vint, f2, hist

    def grad_v_squared_ext(self, side, beads, N, dVdr, dr, gext, periodic=None) -> np.ndarray:
        """gradVSquared[M] with the external potential's gradient per bead (gext shaped like beads)."""
        beads, M, Next, nd = self._beads(beads)
        side, per = self._box(side, periodic)
        dVdr, gext = _f64(dVdr), _f64(gext)
        ext = np.zeros(2)
        f2 = np.zeros(M)
        rc = self.lib.orc_grad_v_squared_ext(nd, _dptr(side), per.ctypes.data_as(_up), _dptr(beads), M, N, Next, _dptr(dVdr),
                                             len(dVdr), dr, _dptr(ext), _dptr(gext), _dptr(f2))
        assert rc == 0
        return f2

    def potential_action(self, vint, f2, VFactor, gradVFactor, tau, lam) -> float:
        vint = _f64(vint)
        f2 = _f64(f2) if f2 is not None else np.zeros_like(vint)
        vf, gf = _f64(VFactor), _f64(gradVFactor)
        return self.lib.orc_potential_action(_dptr(vint), _dptr(f2), len(vint), _dptr(vf), _dptr(gf), tau, lam)

    def deriv_potential_action_tau(self, vint_s, f2_s, slice_, VFactor, gradVFactor, tau, lam) -> float:
        vf, gf = _f64(VFactor), _f64(gradVFactor)
        return self.lib.orc_deriv_potential_action_tau(vint_s, f2_s, slice_, _dptr(vf), _dptr(gf), tau, lam)

    def deriv_potential_action_lambda(self, f2_s, slice_, gradVFactor, tau) -> float:
        gf = _f64(gradVFactor)
        return self.lib.orc_deriv_potential_action_lambda(f2_s, slice_, _dptr(gf), tau)

    def aziz_tail(self, rc, year=1979) -> float:
        return self.lib.orc_aziz_tail(year, rc)

    def energy(self, side, beads, N, vint, f2, VFactor, gradVFactor, period, tau, lam, tailV, mu=0.0, next_links=None) -> np.ndarray:
        """EnergyEstimator::accumulate for one configuration -> [K, V, V_ext, V_int, E, E_mu, K/N, V/N, E/N]."""
        beads, M, Next, nd = self._beads(beads)
        side = _f64(side)
        per = np.ones(nd, dtype=np.uint32)
        vint, f2 = _f64(vint), _f64(f2)
        vf, gf = _f64(VFactor), _f64(gradVFactor)
        nl = None
        if next_links is not None:
            nl = np.ascontiguousarray(next_links, dtype=np.int32)
        out = np.zeros(9)
        rc = self.lib.orc_energy(nd, _dptr(side), per.ctypes.data_as(_up), _dptr(beads), M, N, Next,
                                 nl.ctypes.data_as(C.POINTER(C.c_int)) if nl is not None else None, _dptr(vint), _dptr(f2),
                                 _dptr(vf), _dptr(gf), period, tau, lam, mu, tailV, _dptr(out))
        assert rc == 0
        return out

    # -- scattering variants ---------------------------------------------------------
    def qvectors2(self, ndim, dq, qmax, geometry="line"):
        """getQVectors2: list of per-magnitude arrays [n_k][ndim] (shell 0 = the null vector)."""
        nq = C.c_int(0)
        r = self.lib.orc_qvectors2(ndim, dq, qmax, geometry.encode(), None, 0, None, 0, C.byref(nq))
        if r == -1000:
            raise ValueError("geometry must be 'line' or 'sphere'")
        nshell = -r if r < 0 else r
        nshell = max(nshell, 1)
        # first pass returned -max(numq, nshell); count shells by a second sizing pass with generous capacity
        out = np.zeros((max(nq.value, 1), ndim))
        sizes = np.zeros(max(nq.value, 1) + 8, dtype=np.int32)
        r = self.lib.orc_qvectors2(ndim, dq, qmax, geometry.encode(), _dptr(out), len(out), _iptr(sizes), len(sizes), C.byref(nq))
        assert r > 0, r
        shells, k = [], 0
        for n in sizes[:r]:
            shells.append(out[k:k + n].copy())
            k += n
        return shells

    def ssf_cyl(self, side, beads, N, q, maxR, periodic=None):
        """Cylinder S(q) raw sums per wave-vector + num1DParticles (slice 0)."""
        beads, M, Next, nd = self._beads(beads)
        side, per = self._box(side, periodic)
        q = _f64(q)
        out = np.zeros(len(q))
        n_in = C.c_int(0)
        rc = self.lib.orc_ssf_cyl(nd, _dptr(side), per.ctypes.data_as(_up), _dptr(beads), M, N, Next, _dptr(q), len(q),
                                  maxR, _dptr(out), C.byref(n_in))
        assert rc == 0
        return out, n_in.value

    def elastic(self, beads, N, q, nthreads=1) -> np.ndarray:
        beads, M, Next, nd = self._beads(beads)
        q = _f64(q)
        out = np.zeros(len(q))
        rc = self.lib.orc_elastic(nd, _dptr(beads), M, N, Next, _dptr(q), len(q), _dptr(out), nthreads)
        assert rc == 0
        return out

    # -- virial ------------------------------------------------------------------------
    @staticmethod
    def _links(next_links):
        return np.ascontiguousarray(next_links, dtype=np.int32) if next_links is not None else None

    def virial_delta(self, side, beads, N, window, next_links=None, periodic=None) -> np.ndarray:
        beads, M, Next, nd = self._beads(beads)
        side, per = self._box(side, periodic)
        nl = self._links(next_links)
        out = np.zeros_like(beads)
        rc = self.lib.orc_virial_delta(nd, _dptr(side), per.ctypes.data_as(_up), _dptr(beads), M, N, Next, _iptr(nl), window, _dptr(out))
        assert rc == 0
        return out

    def virial_sums(self, side, beads, N, window, dVdr, d2V, dr, t2_parity=-1, next_links=None, periodic=None, nthreads=1,
                    gext=None, g2ext=None):
        """[M][4] = {sum gV.r, sum (gV T).r, sum gV.delta, sum (gV T).delta} per slice.  gext [M][Next][nd] / g2ext [M][Next]:
        gradient and Laplacian of the external potential per bead (None = free)."""
        beads, M, Next, nd = self._beads(beads)
        side, per = self._box(side, periodic)
        nl = self._links(next_links)
        dVdr, d2V = _f64(dVdr), _f64(d2V)
        ext = np.zeros(2)
        out = np.zeros((M, 4))
        ge = _f64(gext) if gext is not None else None
        g2 = _f64(g2ext) if g2ext is not None else None
        assert ge is None or ge.shape == beads.shape
        assert g2 is None or g2.shape == beads.shape[:2]
        rc = self.lib.orc_virial_sums_ext(nd, _dptr(side), per.ctypes.data_as(_up), _dptr(beads), M, N, Next, _iptr(nl), window,
                                          _dptr(dVdr), _dptr(d2V), len(dVdr), dr, _dptr(ext), _dptr(ext), t2_parity, _dptr(out), nthreads,
                                          _dptr(ge) if ge is not None else None, _dptr(g2) if g2 is not None else None)
        assert rc == 0
        return out

    VIRIAL_COLUMNS = ("K_op", "K_cv", "V_op", "V_cv", "E", "E_mu", "K_op/N", "K_cv/N", "V_op/N", "V_cv/N", "E/N",
                      "EEcv*Beta^2", "Ecv*Beta", "dEdB", "CvCov1", "CvCov2", "CvCov3", "E_th", "P")

    def virial_energy(self, side, beads, N, window, vir, vint, f2, VFactor, gradVFactor, tau, lam, tailV, mu=0.0,
                      next_links=None, quirk=True) -> np.ndarray:
        """VirialEnergyEstimator::accumulate for one configuration -> the 19 columns of VIRIAL_COLUMNS."""
        beads, M, Next, nd = self._beads(beads)
        side, per = self._box(side, None)
        nl = self._links(next_links)
        vir, vint, f2 = _f64(vir), _f64(vint), _f64(f2)
        vf, gf = _f64(VFactor), _f64(gradVFactor)
        out = np.zeros(19)
        rc = self.lib.orc_virial_energy(nd, _dptr(side), per.ctypes.data_as(_up), _dptr(beads), M, N, Next, _iptr(nl), window,
                                        _dptr(vir), _dptr(vint), _dptr(f2), _dptr(vf), _dptr(gf), tau, lam, mu, tailV,
                                        int(quirk), _dptr(out))
        assert rc == 0
        return out

    # -- output formatting -----------------------------------------------------------
    def format_row(self, estimator, norm, num_accumulated) -> str:
        est, nrm = _f64(estimator), _f64(norm)
        buf = C.create_string_buffer(16 * len(est) + 8)
        n = self.lib.orc_format_row(_dptr(est), _dptr(nrm), len(est), num_accumulated, buf, len(buf))
        assert n >= 0
        return buf.value.decode()

    def dvec_to_string(self, v) -> str:
        v = _f64(v)
        buf = C.create_string_buffer(64 * len(v))
        self.lib.orc_dvec_to_string(len(v), _dptr(v), buf, len(buf))
        return buf.value.decode()


_cache = {}


def get(fast: bool = False) -> Oracle:
    if fast not in _cache:
        _cache[fast] = Oracle(fast)
    return _cache[fast]
