"""Parity against the REFERENCE'S OWN code: src/estimator_gpu.cu of the upstream tree, compiled unmodified into
oracle/_ref/librefgpu<NDIM>d.so (oracle/Makefile, target `ref`; built by __graft_entry__.build() wherever
/root/reference is present, shipped to the GPU box with the snapshot).  Its kernels are the upstream GPU estimators
"static structure factor gpu" / "intermediate scattering function gpu" / "elastic scattering gpu":

    ssf_ref[q]      = (2/N)    sum_t sum_{i,j} cos(q.(r_j - r_i))                     = 2 * (sf/N of the CPU estimator)
    isf_ref[q][tau] = (2/(N M)) sum_t sum_{i,j} cos(q.(r_j(t+tau) - r_i(t))), tau <= M/2 = 2/M * (isf/N of the CPU estimator)
    es_ref[q]       = sum_{tau <= M/2} isf_ref[q][tau]                                   (atomicAdd of the tau blocks)

(raw positions; for wave-vectors commensurate with the box this is the CPU estimator's minimum-image S(q) as well).
Both our CUDA path and the CPU restatement (oracle/) are held to 1e-10 against it.
"""
import ctypes as C
import os

import numpy as np
import pytest

from parity import assert_parity
from pimc_b200 import synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_dp = C.POINTER(C.c_double)


def ref_lib(ndim):
    path = os.path.join(ROOT, "oracle", "_ref", f"librefgpu{ndim}d.so")
    if not os.path.exists(path):
        pytest.skip(f"{path} not built (needs the upstream tree: make -C oracle ref)")
    lib = C.CDLL(path)
    for f in (lib.ref_gpu_ssf, lib.ref_gpu_isf, lib.ref_gpu_es):
        f.argtypes = [_dp, C.c_int, C.c_int, C.c_int, _dp, C.c_int, _dp]
    assert lib.ref_ndim() == ndim
    return lib


def reference(lib, beads, N, q):
    beads = np.ascontiguousarray(beads, dtype=np.float64)
    q = np.ascontiguousarray(q, dtype=np.float64)
    M, Next, _ = beads.shape
    ssf = np.zeros(len(q))
    isf = np.zeros((len(q), M // 2 + 1))
    assert lib.ref_gpu_ssf(beads.ctypes.data_as(_dp), M, N, Next, q.ctypes.data_as(_dp), len(q), ssf.ctypes.data_as(_dp)) == 0
    assert lib.ref_gpu_isf(beads.ctypes.data_as(_dp), M, N, Next, q.ctypes.data_as(_dp), len(q), isf.ctypes.data_as(_dp)) == 0
    return ssf, isf


def reference_es(lib, beads, N, q):
    beads = np.ascontiguousarray(beads, dtype=np.float64)
    q = np.ascontiguousarray(q, dtype=np.float64)
    M, Next, _ = beads.shape
    es = np.zeros(len(q))
    assert lib.ref_gpu_es(beads.ctypes.data_as(_dp), M, N, Next, q.ctypes.data_as(_dp), len(q), es.ctypes.data_as(_dp)) == 0
    return es


@pytest.mark.parametrize("name", ["C1", "C2", "2d", "ragged"])
def test_ours_and_oracle_against_upstream_gpu_kernels(orc, nthreads, name):
    from pimc_b200 import api
    if name == "C1":
        s, seed = synth.C1, 3
        q = orc.qvectors("int", synth.int_wavevector_text(s.nq, 3), s.side)
    elif name == "C2":
        s, seed = synth.C2, 4
        q = synth.commensurate_q(s.nq, s.side)
    elif name == "2d":
        s, seed = synth.Shape("r2", 2, 40, 30, 1.0, 0.0432, 0), 5
        q = orc.qvectors("max_int", "3 3", s.side)
    else:
        s, seed = synth.Shape("rr", 3, 37, 10, 2.0, 0.02198, 0), 6
        q = synth.commensurate_q(11, s.side, include_zero=True)
    beads = synth.gen_config(s.N, s.M, s.ndim, s.rho, s.T, seed=synth.BASE_SEED + seed, pad=3)
    beads[:, s.N:, :] = 4321.0                      # the padding columns must not matter to either implementation
    r_ssf, r_isf = reference(ref_lib(s.ndim), beads, s.N, q)
    half = s.M // 2 + 1
    with api.Context(0, s.ndim) as ctx:
        ctx.set_box(s.side)
        ctx.set_qvecs(q)
        ssf, isf = ctx.stage(beads, s.N).ssf_isf()
        es = ctx.elastic()                      # reuses the pass above
        es_fresh = ctx.stage(beads, s.N).elastic()
    # upstream "elastic scattering gpu" (gpu_isf<true>, atomicAdd of the M/2+1 tau blocks into es[q])
    r_es = reference_es(ref_lib(s.ndim), beads, s.N, q)
    assert_parity(es[0], r_es, f"{name}: our elastic scattering vs upstream gpu_es")
    assert np.array_equal(es, es_fresh)
    assert_parity(2.0 * ssf[0], r_ssf, f"{name}: our S(q) vs upstream gpu_ssf")
    assert_parity(2.0 / s.M * isf[0][:, :half], r_isf, f"{name}: our F(q,tau) vs upstream gpu_isf")
    # the CPU restatement against the same upstream code (the full direct loop where it is affordable)
    assert_parity(2.0 * orc.ssf(s.side, beads, s.N, q, nthreads=nthreads), r_ssf, f"{name}: oracle S(q) vs upstream gpu_ssf")
    if name == "C2":
        o_isf = orc.isf_factorised(beads, s.N, q)
        k = 21
        assert_parity(2.0 / s.M * orc.isf(beads, s.N, q[k:k + 1], nthreads=nthreads)[:, :half], r_isf[k:k + 1],
                      "C2: oracle direct F(q,tau) vs upstream gpu_isf (1 q)")
    else:
        o_isf = orc.isf(beads, s.N, q, nthreads=nthreads)
    assert_parity(2.0 / s.M * o_isf[:, :half], r_isf, f"{name}: oracle F(q,tau) vs upstream gpu_isf")
    if name != "C2":
        assert_parity(orc.elastic(beads, s.N, q, nthreads=nthreads), r_es, f"{name}: oracle elastic scattering vs upstream gpu_es")
