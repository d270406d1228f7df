"""Parity against the REFERENCE'S OWN CPU code for the whole hot path.  The upstream function bodies -- the estimators'
accumulate() loops, the LocalAction slice sums, getQVectors / getQVectors2, putInBC, getSeparation / getVelocity, the Aziz
class -- are cut out of the upstream tree at build time and compiled into oracle/_ref/librefcpu<NDIM>d.so
(oracle/ref_cpu_extract.py + oracle/ref_cpu_shim.cpp, `make -C oracle ref`; built by __graft_entry__.build() wherever
/root/reference exists, shipped with the snapshot).  The CPU oracle is held against them here (CPU suite), and the CUDA
path directly (GPU suite) -- to the north-star tolerance of 1e-10, integer outputs exactly."""
import math

import numpy as np
import pytest

from parity import assert_parity
from pimc_b200 import synth
from refcpu import RefCpu

LAM = synth.LAMBDA_HE4
GSF = ([2.0 / 3.0, 4.0 / 3.0], [0.0, 2.0 / 9.0], 2)                # src/setup.cpp:1240-1246
PRIMITIVE = ([1.0, 1.0], [0.0, 0.0], 1)
LI_BROUGHTON = ([1.0, 1.0], [1.0 / 12.0, 1.0 / 12.0], 2)


def permuted_links(M, N, Next, seed):
    rng = np.random.default_rng(seed)
    nxt = np.full((M, Next, 2), -1, dtype=np.int32)
    for s in range(M):
        nxt[s, :N, 0] = (s + 1) % M
        nxt[s, :N, 1] = np.arange(N)
    nxt[M - 1, :N, 1] = rng.permutation(N)
    return nxt


# ------------------------------------------------------------------------------------------------------------------
# CPU: oracle vs upstream code
# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("ndim,N,rho,qtype,text", [
    (3, 16, 0.02198, "int", "1 0 0  0 -2 3  5 5 5"),
    (3, 16, 0.02198, "float", "0.1 0.25 1.7  -2.2 0.3 0.0"),
    (3, 16, 0.02198, "max_int", "2 1 2"),
    (3, 256, 0.02198, "max_float", "0.9 0.0 0.0"),
    (2, 128, 0.0432, "max_int", "8 8"),
    (2, 128, 0.0432, "max_float", "1.1 0.0"),
    (2, 128, 0.0432, "int", "1 0 0 1 -1 1"),
])
def test_qvectors_bit_identical_to_upstream(orc, ndim, N, rho, qtype, text):
    side = np.full(ndim, (N / rho) ** (1.0 / ndim))
    ref = RefCpu(ndim).qvectors(qtype, text, side)
    got = orc.qvectors(qtype, text, side)
    assert got.shape == ref.shape and np.array_equal(got, ref)


@pytest.mark.parametrize("ndim,geometry,qmax", [(3, "line", 4.0), (3, "sphere", 0.9), (2, "line", 2.0), (2, "sphere", 1.0)])
def test_qvectors2_bit_identical_to_upstream(orc, ndim, geometry, qmax):
    side = np.full(ndim, 20.0)
    dq = 2.0 * math.pi / side[-1]
    ref = RefCpu(ndim).qvectors2(dq, qmax, geometry, side)
    got = orc.qvectors2(ndim, dq, qmax, geometry)
    assert len(got) == len(ref)
    for a, b in zip(got, ref):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("name", ["C1", "2d", "ragged", "noncommensurate", "slab"])
def test_ssf_isf_oracle_vs_upstream_loops(orc, nthreads, name):
    """StaticStructureFactorEstimator::accumulate (minimum image) and IntermediateScatteringFunctionEstimator::accumulate
    (raw positions): the restatement reproduces the upstream loops bit for bit (same operation order, same flags)."""
    periodic = None
    if name == "C1":
        s, q = synth.C1, synth.commensurate_q(64, synth.C1.side)
    elif name == "2d":
        s = synth.Shape("r2", 2, 40, 30, 1.0, 0.0432, 0)
        q = orc.qvectors("max_int", "3 3", s.side)
    elif name == "ragged":
        s = synth.Shape("rr", 3, 37, 10, 2.0, 0.02198, 0)
        q = synth.commensurate_q(11, s.side, include_zero=True)
    elif name == "noncommensurate":
        s = synth.Shape("nc", 3, 20, 8, 2.0, 0.02198, 0)
        q = synth.float_q(9, 3)
    else:
        s = synth.Shape("sl", 3, 20, 8, 2.0, 0.02198, 0)
        q = np.vstack([synth.commensurate_q(5, s.side), synth.float_q(4, 3)])
        periodic = [1, 1, 0]
    beads = synth.gen_config(s.N, s.M, s.ndim, s.rho, s.T, seed=synth.BASE_SEED + 9, pad=3)
    beads[:, s.N:, :] = 4321.0
    ref = RefCpu(s.ndim)
    r_ssf = ref.ssf(s.side, beads, s.N, q, periodic)
    o_ssf = orc.ssf(s.side, beads, s.N, q, periodic=periodic, nthreads=nthreads)
    assert np.array_equal(o_ssf, r_ssf)
    nq_isf = min(len(q), 6)
    r_isf = ref.isf(s.side, beads, s.N, q[:nq_isf])
    o_isf = orc.isf(beads, s.N, q[:nq_isf], nthreads=nthreads)
    assert np.array_equal(o_isf, r_isf)
    assert_parity(orc.isf_factorised(beads, s.N, q[:nq_isf]), r_isf, f"{name}: factorised CPU variant vs upstream loop")


def test_cylinder_ssf_oracle_vs_upstream(orc):
    from test_variants import cylinder_config
    N, M, L, R = 37, 12, 11.0, 3.0
    beads, side, per = cylinder_config(N, M, L, R, seed=3)
    ref = RefCpu(3)
    shells = ref.qvectors2(2.0 * math.pi / L, 4.0, "line", side)
    maxR = 2.0
    r_out, n1d = ref.ssf_cyl(side, beads, N, shells, maxR, per)
    raw, n_in = orc.ssf_cyl(side, beads, N, np.vstack(shells), maxR, periodic=per)
    assert n_in == n1d and 0 < n1d < N
    assert np.array_equal(raw / n_in, r_out)                       # one vector per shell for "line"
    # several vectors per shell ("sphere"): shell sums
    sph = ref.qvectors2(2.0 * math.pi / L, 1.2, "sphere", side)
    r_out, n1d = ref.ssf_cyl(side, beads, N, sph, maxR, per)
    raw, _ = orc.ssf_cyl(side, beads, N, np.vstack(sph), maxR, periodic=per)
    k, sums = 0, []
    for sh in sph:
        sums.append(raw[k:k + len(sh)].sum())
        k += len(sh)
    np.testing.assert_allclose(np.array(sums) / n1d, r_out, rtol=1e-13)


@pytest.mark.parametrize("action", ["gsf", "primitive", "li_broughton"])
def test_action_sums_oracle_vs_upstream(orc, nthreads, action):
    """LocalAction::V(slice) + sepHist, gradVSquared, potentialAction and its derivatives, the virial terms, and the
    energy / virial estimators: oracle restatement vs the upstream bodies running on the upstream Aziz class."""
    VF, GF, period = {"gsf": GSF, "primitive": PRIMITIVE, "li_broughton": LI_BROUGHTON}[action]
    s = synth.Shape("ac", 3, 18, 8, 2.0, 0.02198, 0)
    beads = synth.gen_config(s.N, s.M, 3, s.rho, s.T, seed=77, pad=2)
    Next = beads.shape[1]
    links = permuted_links(s.M, s.N, Next, 4)
    window = 3
    r = RefCpu(3).action(s.side, beads, s.N, s.tau, LAM, VF, GF, period, window=window, mu=-1.5, next_links=links)
    V, dV, d2V, dr = orc.aziz_table(orc.max_sep(s.side), second=True)
    dSep = 0.5 * math.sqrt(3) * s.side[2] / 50
    vint, f2, hist = orc.pair_sums(s.side, beads, s.N, V, dV, dr, dSep, nthreads=nthreads)
    assert np.array_equal(hist, r["sephist"])
    assert np.array_equal(vint, r["vint"])
    assert np.array_equal(f2, r["f2"])
    f2m = f2.copy()
    for eo in (0, 1):
        if not GF[eo] > 1e-7:
            f2m[eo::2] = 0.0
    np.testing.assert_allclose(orc.potential_action(vint, f2m, VF, GF, s.tau, LAM), r["potentialAction"], rtol=1e-14)
    for t in range(s.M):
        np.testing.assert_allclose(orc.deriv_potential_action_tau(vint[t], f2m[t], t, VF, GF, s.tau, LAM), r["dtau"][t], rtol=1e-14)
        np.testing.assert_allclose(orc.deriv_potential_action_lambda(f2m[t], t, GF, s.tau), r["dlam"][t], rtol=1e-14, atol=0)
    # virial terms: the oracle returns the bare sums, upstream multiplies by VFactor*tau and 2*gradVFactor*tau^3*lambda
    t2p = -1 if (GF[0] > 1e-7 and GF[1] > 1e-7) else (1 if GF[1] > 1e-7 else (0 if GF[0] > 1e-7 else -2))
    vir = orc.virial_sums(s.side, beads, s.N, window, dV, d2V, dr, t2_parity=t2p, next_links=links, nthreads=nthreads)
    eo = np.arange(s.M) % 2
    c1 = np.array(VF)[eo] * s.tau
    c2 = 2.0 * np.array(GF)[eo] * s.tau ** 3 * LAM
    assert_parity(vir[:, 0] * c1, r["vir"][:, 0], "rDOTgradUterm1")
    assert_parity(vir[:, 2] * c1, r["vir"][:, 2], "deltaDOTgradUterm1")
    if t2p != -2:
        assert_parity(vir[:, 1] * c2, r["vir"][:, 1], "rDOTgradUterm2")
        assert_parity(vir[:, 3] * c2, r["vir"][:, 3], "deltaDOTgradUterm2")
    else:
        assert np.all(r["vir"][:, 1] == 0.0) and np.all(r["vir"][:, 3] == 0.0)
    tail = orc.aziz_tail(s.side[2])
    en = orc.energy(s.side, beads, s.N, vint, f2m, VF, GF, period, s.tau, LAM, tail, mu=-1.5, next_links=links)
    np.testing.assert_allclose(en, r["energy"], rtol=1e-12, atol=1e-12 * np.max(np.abs(r["energy"])))
    ve = orc.virial_energy(s.side, beads, s.N, window, vir, vint, f2m, VF, GF, s.tau, LAM, tail, mu=-1.5, next_links=links)
    np.testing.assert_allclose(ve, r["virial"], rtol=1e-11, atol=1e-11 * np.max(np.abs(r["virial"])))
    assert r["virial"][15] == 0.0                                   # the upstream "cVCov2" key typo: the column stays empty


def test_external_gradient_coupling_oracle_vs_upstream(orc):
    """gradVSquared with a non-trivial external potential (F_i += gradVext(r_i) before squaring, src/action.cpp:1216):
    oracle vs the upstream body running with a spring potential V = k r^2/2."""
    VF, GF, period = LI_BROUGHTON
    s = synth.Shape("xg", 3, 14, 6, 2.0, 0.02198, 0)
    beads = synth.gen_config(s.N, s.M, 3, s.rho, s.T, seed=41, pad=2)
    k = 40.0
    r = RefCpu(3).action(s.side, beads, s.N, s.tau, LAM, VF, GF, period, spring_k=k)
    V, dV, dr = orc.aziz_table(orc.max_sep(s.side))
    gext = k * beads
    assert np.array_equal(orc.grad_v_squared_ext(s.side, beads, s.N, dV, dr, gext), r["f2"])
    free = RefCpu(3).action(s.side, beads, s.N, s.tau, LAM, VF, GF, period)
    assert not np.allclose(free["f2"], r["f2"], rtol=1e-3)            # the coupling matters
    np.testing.assert_allclose(orc.grad_v_squared_ext(s.side, beads, s.N, dV, dr, 0.0 * beads), free["f2"], rtol=0, atol=0)


def spring_terms(beads, k):
    """SpringTestPotential of oracle/ref_cpu_shim.cpp: V = k r^2 / 2 -> gradV = k r, grad2V = NDIM k."""
    return k * beads, np.full(beads.shape[:2], beads.shape[2] * k)


@pytest.mark.parametrize("action", ["gsf", "li_broughton", "primitive"])
def test_virial_terms_with_external_potential_oracle_vs_upstream(orc, nthreads, action):
    """The four virial terms with a non-free external potential (src/action.cpp:1446-1784): gV = gVe + sum gVi in the
    first-order terms, dV = dVi + dVe and d2V = g2Vi + g2Ve inside the T-matrix of the second-order ones.  Oracle
    restatement vs the upstream bodies running with a spring potential, permuted world lines."""
    VF, GF, period = {"gsf": GSF, "primitive": PRIMITIVE, "li_broughton": LI_BROUGHTON}[action]
    s = synth.Shape("xv", 3, 15, 6, 2.0, 0.02198, 0)
    beads = synth.gen_config(s.N, s.M, 3, s.rho, s.T, seed=91, pad=2)
    links = permuted_links(s.M, s.N, beads.shape[1], 5)
    window, k = 2, 40.0
    r = RefCpu(3).action(s.side, beads, s.N, s.tau, LAM, VF, GF, period, window=window, next_links=links, spring_k=k)
    free = RefCpu(3).action(s.side, beads, s.N, s.tau, LAM, VF, GF, period, window=window, next_links=links)
    V, dV, d2V, dr = orc.aziz_table(orc.max_sep(s.side), second=True)
    t2p = -1 if (GF[0] > 1e-7 and GF[1] > 1e-7) else (1 if GF[1] > 1e-7 else (0 if GF[0] > 1e-7 else -2))
    gext, g2ext = spring_terms(beads, k)
    vir = orc.virial_sums(s.side, beads, s.N, window, dV, d2V, dr, t2_parity=t2p, next_links=links, nthreads=nthreads,
                          gext=gext, g2ext=g2ext)
    eo = np.arange(s.M) % 2
    c1 = np.array(VF)[eo] * s.tau
    c2 = 2.0 * np.array(GF)[eo] * s.tau ** 3 * LAM
    np.testing.assert_allclose(vir[:, 0] * c1, r["vir"][:, 0], rtol=1e-13)
    np.testing.assert_allclose(vir[:, 2] * c1, r["vir"][:, 2], rtol=1e-12, atol=1e-13 * np.max(np.abs(r["vir"][:, 2])))
    assert not np.allclose(free["vir"][:, 0], r["vir"][:, 0], rtol=1e-3)          # the external force matters
    if t2p != -2:
        np.testing.assert_allclose(vir[:, 1] * c2, r["vir"][:, 1], rtol=1e-12, atol=1e-13 * np.max(np.abs(r["vir"][:, 1])))
        np.testing.assert_allclose(vir[:, 3] * c2, r["vir"][:, 3], rtol=1e-12, atol=1e-13 * np.max(np.abs(r["vir"][:, 3])))
        on = c2 > 0
        assert not np.allclose(free["vir"][on, 1], r["vir"][on, 1], rtol=1e-5)
    # and with no external arrays the restatement is unchanged
    np.testing.assert_array_equal(orc.virial_sums(s.side, beads, s.N, window, dV, d2V, dr, t2_parity=t2p, next_links=links),
                                  orc.virial_sums(s.side, beads, s.N, window, dV, d2V, dr, t2_parity=t2p, next_links=links,
                                                  gext=0.0 * beads, g2ext=np.zeros(beads.shape[:2])))


# ------------------------------------------------------------------------------------------------------------------
# GPU: CUDA path vs upstream CPU code, directly
# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_cuda_path_against_upstream_cpu_code(orc):
    from pimc_b200 import api
    s = synth.C1
    beads = synth.gen_config(s.N, s.M, 3, s.rho, s.T, seed=synth.BASE_SEED + 31, pad=3)
    Next = beads.shape[1]
    ref = RefCpu(3)
    text = synth.int_wavevector_text(24, 3)
    q = np.vstack([ref.qvectors("int", text, s.side), ref.qvectors("float", "0.3 -1.1 0.7 2.0 0.1 0.4", s.side)])
    links = permuted_links(s.M, s.N, Next, 12)
    VF, GF, period = GSF
    window = 5
    r = ref.action(s.side, beads, s.N, s.tau, LAM, VF, GF, period, window=window, next_links=links)
    V, dV, d2V, dr = orc.aziz_table(orc.max_sep(s.side), second=True)       # array_equal to upstream (test_reference_aziz.py)
    dSep = 0.5 * math.sqrt(3) * s.side[2] / 50
    delta = orc.virial_delta(s.side, beads, s.N, window, next_links=links)
    with api.Context(0, 3) as ctx:
        ctx.set_box(s.side)
        ctx.set_qvecs(q)
        ctx.set_pair_table(V, dV, dr)
        ctx.set_pair_table_d2(d2V)
        ssf, isf = ctx.stage(beads, s.N).ssf_isf()
        vint, f2, hist = ctx.pair_sums(dSep, f2_parity=-1)
        vir = ctx.virial_sums(delta, t2_parity=1)
    assert_parity(ssf[0], ref.ssf(s.side, beads, s.N, q), "S(q) vs upstream CPU loop (min-image, incl. non-commensurate q)")
    k = [0, 7, 23, 25]
    assert_parity(isf[0][k], ref.isf(s.side, beads, s.N, q[k]), "F(q,tau) vs upstream CPU loop")
    assert np.array_equal(hist[0], r["sephist"])
    assert_parity(vint[0], r["vint"], "Vint vs upstream LocalAction::V(slice)")
    assert_parity(f2[0], r["f2"], "gradVSquared vs upstream")
    eo = np.arange(s.M) % 2
    c1 = np.array(VF)[eo] * s.tau
    c2 = 2.0 * np.array(GF)[eo] * s.tau ** 3 * LAM
    assert_parity(vir[0][:, 0] * c1, r["vir"][:, 0], "rDOTgradUterm1 vs upstream")
    assert_parity(vir[0][:, 1] * c2, r["vir"][:, 1], "rDOTgradUterm2 vs upstream")
    assert_parity(vir[0][:, 2] * c1, r["vir"][:, 2], "deltaDOTgradUterm1 vs upstream")
    assert_parity(vir[0][:, 3] * c2, r["vir"][:, 3], "deltaDOTgradUterm2 vs upstream")


@pytest.mark.gpu
def test_cuda_cylinder_ssf_against_upstream_cpu_code(orc):
    from pimc_b200 import api
    from test_variants import cylinder_config
    N, M, L, R = 37, 12, 11.0, 3.0
    beads, side, per = cylinder_config(N, M, L, R, seed=8)
    ref = RefCpu(3)
    shells = ref.qvectors2(2.0 * math.pi / L, 4.0, "line", side)
    r_out, n1d = ref.ssf_cyl(side, beads, N, shells, 2.0, per)
    with api.Context(0, 3) as ctx:
        ctx.set_box(side, per)
        ctx.set_qvecs(np.vstack(shells))
        out, n_in = ctx.stage(beads, N).ssf_cyl(2.0)
    assert n_in[0] == n1d
    assert_parity(out[0] / n_in[0], r_out, "cylinder S(q) vs upstream CPU loop")


@pytest.mark.gpu
def test_cuda_external_gradient_against_upstream_cpu_code(orc):
    from pimc_b200 import api
    VF, GF, period = LI_BROUGHTON
    s = synth.Shape("xg", 3, 37, 10, 2.0, 0.02198, 0)
    batch = synth.gen_batch(s, 2, first=44, pad=2)
    k = 40.0
    V, dV, dr = orc.aziz_table(orc.max_sep(s.side))
    dSep = 0.5 * math.sqrt(3) * s.side[2] / 50
    with api.Context(0, 3) as ctx:
        ctx.set_box(s.side)
        ctx.set_pair_table(V, dV, dr)
        ctx.stage(batch, s.N)
        _, f2_free, _ = ctx.pair_sums(dSep)
        ctx.set_external_gradient(k * batch)
        vint, f2, hist = ctx.pair_sums(dSep)
        ctx.set_external_gradient(None)
        _, f2_cleared, _ = ctx.pair_sums(dSep)
        ctx.set_external_gradient(k * batch)
        ctx.stage(batch[::-1].copy(), s.N)                      # new beads: the old gradient must not be applied to them
        _, f2_restaged, _ = ctx.pair_sums(dSep)
    for b in range(2):
        r = RefCpu(3).action(s.side, batch[b], s.N, s.tau, LAM, VF, GF, period, spring_k=k)
        assert_parity(f2[b], r["f2"], "gradVSquared with external gradient vs upstream")
        assert_parity(vint[b], r["vint"], "Vint")
        assert np.array_equal(hist[b], r["sephist"])
    assert np.array_equal(f2_free, f2_cleared) and np.array_equal(f2_restaged, f2_free[::-1])
    assert not np.allclose(f2_free, f2, rtol=1e-3)


@pytest.mark.gpu
@pytest.mark.parametrize("action", ["gsf", "li_broughton", "primitive"])
def test_cuda_virial_terms_with_external_potential_against_upstream_cpu_code(orc, action):
    """pimcb_virial_sums with the external gradient and Laplacian uploaded (pimcb_set_external_gradient /
    pimcb_set_external_laplacian) vs the upstream rDOTgradUterm1/2 and deltadotgradUterm1/2 bodies running with a
    spring potential; permuted world lines; the arrays are bound to the staged configuration."""
    from pimc_b200 import api
    VF, GF, period = {"gsf": GSF, "primitive": PRIMITIVE, "li_broughton": LI_BROUGHTON}[action]
    s = synth.Shape("xv", 3, 37, 8, 2.0, 0.02198, 0)
    beads = synth.gen_config(s.N, s.M, 3, s.rho, s.T, seed=93, pad=3)
    links = permuted_links(s.M, s.N, beads.shape[1], 6)
    window, k = 3, 40.0
    r = RefCpu(3).action(s.side, beads, s.N, s.tau, LAM, VF, GF, period, window=window, next_links=links, spring_k=k)
    free = RefCpu(3).action(s.side, beads, s.N, s.tau, LAM, VF, GF, period, window=window, next_links=links)
    V, dV, d2V, dr = orc.aziz_table(orc.max_sep(s.side), second=True)
    delta = orc.virial_delta(s.side, beads, s.N, window, next_links=links)
    t2p = -1 if (GF[0] > 1e-7 and GF[1] > 1e-7) else (1 if GF[1] > 1e-7 else (0 if GF[0] > 1e-7 else -2))
    gext, g2ext = spring_terms(beads, k)
    with api.Context(0, 3) as ctx:
        ctx.set_box(s.side)
        ctx.set_pair_table(V, dV, dr)
        ctx.set_pair_table_d2(d2V)
        ctx.stage(beads, s.N)
        vir_free = ctx.virial_sums(delta, t2_parity=t2p)[0]
        ctx.set_external_gradient(gext)
        ctx.set_external_laplacian(g2ext)
        vir = ctx.virial_sums(delta, t2_parity=t2p)[0]
        ctx.stage(beads, s.N)                                   # re-staged: the external arrays are not applied any more
        vir_restaged = ctx.virial_sums(delta, t2_parity=t2p)[0]
    eo = np.arange(s.M) % 2
    c1 = np.array(VF)[eo] * s.tau
    c2 = 2.0 * np.array(GF)[eo] * s.tau ** 3 * LAM
    for got, ref in ((vir, r), (vir_free, free), (vir_restaged, free)):
        assert_parity(got[:, 0] * c1, ref["vir"][:, 0], "rDOTgradUterm1")
        assert_parity(got[:, 2] * c1, ref["vir"][:, 2], "deltaDOTgradUterm1")
        if t2p != -2:
            assert_parity(got[:, 1] * c2, ref["vir"][:, 1], "rDOTgradUterm2")
            assert_parity(got[:, 3] * c2, ref["vir"][:, 3], "deltaDOTgradUterm2")
    assert not np.allclose(vir[:, 0], vir_free[:, 0], rtol=1e-3)


@pytest.mark.parametrize("ndim", [2, 3])
def test_state_file_arrays_against_upstream_stream_operators(ndim, tmp_path):
    """The text format of the saved state (src/pimc.cpp:955-962) is what the upstream operator<< / operator>> of
    std::array and DynamicArray<T,2> (include/common.h:244-399) produce and accept.  The Python restatement of the writer
    (oracle/statefile.py) emits exactly the upstream text for beads, links and worm.beads; the upstream operator>> and
    both our readers (oracle/statefile.py, pimc_b200/host/state_file.cpp through pimcb_host_selftest) recover the same
    arrays bit for bit from it."""
    import os
    import subprocess
    from oracle import statefile
    ref = RefCpu(ndim)
    rng = np.random.default_rng(ndim)
    M, W, N = 5, 6, 4
    beads = rng.uniform(-7.0, 7.0, size=(M, W, ndim))
    beads[0, 0] = 0.0
    beads[1, 1, 0] = 1.0e-5                                # short and exponent forms of operator<<(double)
    beads[2, 2, 0] = -12.3456789125
    on = np.zeros((M, W), dtype=np.uint32)
    on[:, :N] = 1
    nxt = np.full((M, W, 2), -1, dtype=np.int32)
    for t in range(M):
        nxt[t, :N, 0] = (t + 1) % M
        nxt[t, :N, 1] = rng.permutation(N) if t == M - 1 else np.arange(N)
    f = tmp_path / "ce-state-x.dat"
    statefile.write_state(f, beads, on, next_link=nxt)
    text = open(f).read()
    t_beads, t_next, t_worm = ref.write_array(0, beads), ref.write_array(1, nxt), ref.write_array(2, on)
    for piece in (t_beads, t_next, t_worm):
        assert piece in text, "the Python writer must emit the upstream operator<< text verbatim"
    assert text.index(t_beads) < text.index(t_next) < text.index(t_worm)
    # upstream operator>> on the upstream text: exact round trip at setprecision(16)? 17 digits are needed in general, so
    # compare all readers on the SAME text instead
    up_beads = ref.read_array(0, t_beads)
    up_next = ref.read_array(1, t_next)
    up_worm = ref.read_array(2, t_worm)
    assert np.array_equal(up_next, nxt) and np.array_equal(up_worm, on)
    np.testing.assert_allclose(up_beads, beads, rtol=2e-16 * 10)
    mine = statefile.read_state(f)
    assert np.array_equal(mine["beads"], up_beads)
    # C++ parser of the host layer
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "pimc_b200", "host", f"pimcb_host_selftest{ndim}d")
    if not os.path.exists(exe):
        subprocess.run(["make", "-C", os.path.dirname(exe), f"NDIM={ndim}"], check=True, stdout=subprocess.DEVNULL)
    rho = N / 30.0 ** ndim                                  # a cell of side 30: putInside leaves these positions alone
    out = subprocess.run([exe, "--state", str(f), str(N), repr(rho)], check=True, capture_output=True, text=True).stdout
    got = np.array([[float(x) for x in l.split("=", 1)[1].split()] for l in out.splitlines() if l.startswith("bead=")])
    assert np.array_equal(got.reshape(M, N, ndim), up_beads[:, :N])


def test_extraction_recipe_finds_every_upstream_definition(tmp_path):
    """The build-time recipes locate each upstream definition by its signature; a silent mismatch would compile the wrong
    body.  With the upstream tree present: every manifest pattern matches exactly one line, and every cut ends on the brace
    that closes it (balanced braces in the cut text)."""
    import importlib.util
    import os
    import re
    ref = "/root/reference"
    if not os.path.exists(os.path.join(ref, "src", "estimator.cpp")):
        pytest.skip("upstream tree not present")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("ref_cpu_extract", os.path.join(root, "oracle", "ref_cpu_extract.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    for name, (path, pats) in mod.MANIFEST.items():
        lines = open(os.path.join(ref, path)).read().split("\n")
        for pat in pats:
            hits = [i for i, l in enumerate(lines) if re.search(pat, l)]
            assert len(hits) == 1, f"{path}: {pat!r} matches {len(hits)} lines"
            _, body = mod.body(lines, pat)
            depth, in_c = 0, False
            for l in body:
                text, in_c = mod.strip_code(l, in_c)
                depth += text.count("{") - text.count("}")
            assert depth == 0 and len(body) >= 3, f"{path}: {pat!r} cut is unbalanced"
    mod.main(ref, str(tmp_path))
    assert sorted(os.listdir(tmp_path)) == sorted(mod.MANIFEST)


@pytest.mark.parametrize("ndim", [2, 3])
def test_left_pack_against_upstream_body(ndim, tmp_path):
    """Path::leftPack (src/path.cpp:145-189), compiled from the upstream tree (oracle/_ref/librefpack<NDIM>d.so), against
    PimcState::leftPack of the host layer on a hand-made state with holes in every row and permuted world lines: positions,
    flags and BOTH link arrays end up identical (the kinetic and virial estimators follow those links)."""
    import ctypes as C
    import os
    import subprocess
    from oracle import statefile
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    lib_path = os.path.join(root, "oracle", "_ref", f"librefpack{ndim}d.so")
    if not os.path.exists(lib_path):
        pytest.skip(f"{lib_path} not built (needs the upstream tree: make -C oracle ref)")
    lib = C.CDLL(lib_path)
    assert lib.refpack_ndim() == ndim
    rng = np.random.default_rng(10 + ndim)
    M, W, N = 6, 9, 5
    cols = [np.sort(rng.choice(W, size=N, replace=False)) for _ in range(M)]
    on = np.zeros((M, W), dtype=np.uint32)
    beads = np.zeros((M, W, ndim))
    for t in range(M):
        on[t, cols[t]] = 1
        beads[t, cols[t]] = rng.uniform(-3.0, 3.0, size=(N, ndim))
    # closed world lines through the scattered columns, a different permutation on every link
    nxt = np.full((M, W, 2), -1, dtype=np.int32)
    prv = np.full((M, W, 2), -1, dtype=np.int32)
    for t in range(M):
        perm = rng.permutation(N)
        for a in range(N):
            src, dst = (t, cols[t][a]), ((t + 1) % M, cols[(t + 1) % M][perm[a]])
            nxt[src] = dst
            prv[dst] = src
    f = tmp_path / "ce-state-holes.dat"
    statefile.write_state(f, beads, on, next_link=nxt)
    # upstream, on the arrays as the loader sees them (positions at the 16 significant digits of the file)
    parsed = statefile.read_state(f)
    assert np.array_equal(parsed["next"], nxt) and np.array_equal(parsed["prev"], prv)
    u_b, u_n, u_p, u_on = np.ascontiguousarray(parsed["beads"]), nxt.copy(), prv.copy(), on.copy()
    vp = C.c_void_p
    lib.refpack_left_pack.argtypes = [vp, vp, vp, vp, C.c_int, C.c_int]
    assert lib.refpack_left_pack(u_b.ctypes.data, u_n.ctypes.data, u_p.ctypes.data, u_on.ctypes.data, M, W) == 0
    assert np.all(u_on[:, :N] == 1) and np.all(u_on[:, N:] == 0)
    # host layer
    exe = os.path.join(root, "pimc_b200", "host", f"pimcb_host_selftest{ndim}d")
    rho = N / 30.0 ** ndim
    out = subprocess.run([exe, "--state", str(f), str(N), repr(rho)], check=True, capture_output=True, text=True).stdout
    rep = [l.split("=", 1) for l in out.splitlines() if "=" in l]
    assert ["leftPacked", "0"] in rep and ["diagonal", "1"] in rep
    got = np.array([[float(x) for x in v.split()] for k, v in rep if k == "bead"]).reshape(M, N, ndim)
    assert np.array_equal(got, u_b[:, :N])
    links = np.array([[int(x) for x in v.split()] for k, v in rep if k == "link"])
    assert len(links) == M * W
    for s_, p_, ns, np_, ps, pp, flag in links:
        assert flag == u_on[s_, p_]
        assert (ns, np_) == tuple(u_n[s_, p_]) and (ps, pp) == tuple(u_p[s_, p_]), (s_, p_)
