"""Known-answer tests pinning the CPU oracle (the reference ships no fixtures for this path)."""
import math

import numpy as np
import pytest

from pimc_b200 import synth


def lattice_config(n_side, M, L, ndim=3, pad=2):
    a = L / n_side
    g = np.stack(np.meshgrid(*([np.arange(n_side)] * ndim), indexing="ij"), axis=-1).reshape(-1, ndim)
    pos = (g + 0.5) * a - 0.5 * L
    N = len(pos)
    beads = np.zeros((M, N + pad, ndim))
    beads[:, :N, :] = pos[None]
    beads[:, N:, :] = 1234.5     # padding must never be read
    return beads, N


def test_single_particle_S_and_F_are_one(orc):
    side = np.array([7.0, 8.0, 9.0])
    rng = np.random.default_rng(1)
    M = 6
    beads = np.zeros((M, 3, 3))
    beads[:, 0, :] = rng.uniform(-3, 3, size=(M, 3))
    # equal-time: N=1 -> S = 1 for every q; F(q,0) = 1
    q = synth.commensurate_q(5, side)
    ssf = orc.ssf(side, beads, 1, q)
    np.testing.assert_allclose(ssf, M * 1.0, rtol=0, atol=1e-12)
    isf = orc.isf(beads, 1, q)
    np.testing.assert_allclose(isf[:, 0], M * 1.0, rtol=0, atol=1e-12)
    # a static particle has F(q,tau) = 1 for all tau
    beads[:, 0, :] = beads[0, 0, :]
    np.testing.assert_allclose(orc.isf(beads, 1, q), M * 1.0, rtol=0, atol=1e-12)


def test_two_particles_closed_form(orc):
    side = np.array([10.0, 10.0, 10.0])
    M = 4
    beads = np.zeros((M, 2, 3))
    beads[:, 0, :] = [1.0, -2.0, 0.5]
    beads[:, 1, :] = [-3.5, 4.0, 2.0]
    q = synth.float_q(6, 3)
    delta = beads[0, 0] - beads[0, 1]
    delta = delta - side * np.floor(delta / side + 0.5)
    expect = M * (2 + 2 * np.cos(q @ delta)) / 2           # sum_t [N + 2 cos] / N
    np.testing.assert_allclose(orc.ssf(side, beads, 2, q), expect, rtol=1e-14)
    raw = beads[0, 0] - beads[0, 1]                        # ISF uses raw positions
    expect_f = M * (2 + 2 * np.cos(q @ raw)) / 2
    isf = orc.isf(beads, 2, q)
    for tau in range(M):
        np.testing.assert_allclose(isf[:, tau], expect_f, rtol=1e-13)


def test_perfect_lattice_bragg_peaks(orc):
    L, n_side, M = 12.0, 2, 4
    beads, N = lattice_config(n_side, M, L)
    side = np.full(3, L)
    n = synth.lattice_indices(60, 3)
    q = 2 * math.pi / L * n
    ssf = orc.ssf(side, beads, N, q) / M
    bragg = np.all(n % n_side == 0, axis=1)
    assert bragg.sum() >= 6
    np.testing.assert_allclose(ssf[bragg], N, rtol=1e-12)
    np.testing.assert_allclose(ssf[~bragg], 0.0, atol=1e-10)
    isf = orc.isf(beads, N, q, nthreads=4) / M
    np.testing.assert_allclose(isf[bragg], N, rtol=1e-12)
    np.testing.assert_allclose(isf[~bragg], 0.0, atol=1e-10)
    # q = 0 gives S = N
    np.testing.assert_allclose(orc.ssf(side, beads, N, np.zeros((1, 3))) / M, N, rtol=1e-14)


def test_identities_on_synthetic_config(orc):
    s = synth.C1
    beads = synth.gen_config(s.N, s.M, s.ndim, s.rho, s.T)
    q = synth.commensurate_q(12, s.side)
    ssf = orc.ssf(s.side, beads, s.N, q)
    isf = orc.isf(beads, s.N, q, nthreads=4)
    # F(q,0) = S(q) for commensurate q; F(q,tau) = F(q,M-tau)
    np.testing.assert_allclose(isf[:, 0], ssf, rtol=1e-12)
    np.testing.assert_allclose(isf[:, 1:], isf[:, :0:-1], rtol=1e-11, atol=1e-11)
    # factorised variant == direct loop
    np.testing.assert_allclose(orc.isf_factorised(beads, s.N, q), isf, rtol=1e-11, atol=1e-10)
    # relabelling particles and a rigid translation leave S unchanged; F is invariant too
    perm = np.random.default_rng(3).permutation(s.N)
    b2 = beads.copy()
    b2[:, :s.N] = beads[:, perm]
    np.testing.assert_allclose(orc.ssf(s.side, b2, s.N, q), ssf, rtol=1e-12)
    b3 = beads.copy()
    b3[:, :s.N] += np.array([0.3, -1.1, 2.2])
    np.testing.assert_allclose(orc.ssf(s.side, b3, s.N, q), ssf, rtol=1e-11)
    np.testing.assert_allclose(orc.isf_factorised(b3, s.N, q), isf, rtol=1e-10, atol=1e-9)
    # threading over (q,tau) does not change a single bit
    assert np.array_equal(orc.isf(beads, s.N, q[:2], nthreads=1), orc.isf(beads, s.N, q[:2], nthreads=5))


def test_put_in_bc(orc):
    side = np.array([4.0, 6.0])
    r = np.array([[2.0, -3.0], [2.1, 3.1], [-6.3, 0.0], [9.9, -8.9]])
    out = orc.put_in_bc(side, r)
    np.testing.assert_allclose(out, [[-2.0, -3.0], [-1.9, -2.9], [1.7, 0.0], [1.9, -2.9]], atol=1e-12)
    np.testing.assert_array_equal(out, synth.put_in_bc(r, side))
    assert orc.max_sep(side) == pytest.approx(math.sqrt(4 + 9))
    assert orc.max_sep(side, periodic=[1, 0]) == pytest.approx(math.sqrt(4 + 36))


def test_qvector_generation(orc):
    side = np.array([9.0, 10.0, 11.0])
    k = 2 * math.pi / side
    q = orc.qvectors("int", "1 0 0  0 -2 3", side)
    np.testing.assert_allclose(q, [[k[0], 0, 0], [0, -2 * k[1], 3 * k[2]]], rtol=1e-15)
    qf = orc.qvectors("float", "0.1 0.2 0.3", side)
    assert qf[0, 0] == float(np.float32(0.1)) and qf[0, 0] != 0.1     # std::stof rounding
    qm = orc.qvectors("max_int", "1 1 1", side)
    assert qm.shape == (27, 3)
    np.testing.assert_allclose(qm[0], -k)                   # starts at the all-negative corner
    np.testing.assert_allclose(qm[1], [-k[0], -k[1], 0.0], atol=1e-15)   # last dimension fastest
    np.testing.assert_allclose(qm[13], 0.0, atol=1e-15)
    q2 = orc.qvectors("max_int", "8 8", np.array([5.0, 5.0]))
    assert q2.shape == (289, 2)
    # max_float keeps |q| <= |q_max| and produces lattice vectors
    qx = orc.qvectors("max_float", "1.0 0.0 0.0", side)
    assert len(qx) > 0 and np.all(np.linalg.norm(qx, axis=1) <= 1.0 + 1e-12)
    n = qx * side / (2 * math.pi)
    np.testing.assert_allclose(n, np.rint(n), atol=1e-12)
    with pytest.raises(ValueError):
        orc.qvectors("int", "", side)


def test_time_slices(orc):
    assert orc.time_slices(2.0, tau=0.004) == (124, 0.004)          # C1: 125 forced even
    M, tau = orc.time_slices(1.5, P=170)
    assert M == 170 and tau == pytest.approx(1 / (1.5 * 170))
    assert orc.time_slices(1.5, P=171)[0] == 170


def test_aziz_potential(orc):
    rm = orc.aziz_rm(1979)
    assert rm == 2.9673
    r = np.linspace(2.0, 8.0, 4001)
    V = orc.aziz_values(r)
    i = np.argmin(V)
    assert abs(r[i] - rm) < 0.01 and abs(V[i] + 10.8) < 0.01           # well depth eps at r ~ rm
    # independent numpy evaluation of the HFDHE2 form
    eps_, A, al, D, C6, C8, C10 = 10.8, 0.5448504e6, 13.353384, 1.241314, 1.3732412, 0.4253785, 0.1781
    x = r / rm
    F = np.where(x < D, np.exp(-(D / x - 1) ** 2), 1.0)
    ref = eps_ * (A * np.exp(-al * x) - (C6 / x**6 + C8 / x**8 + C10 / x**10) * F)
    np.testing.assert_allclose(V, ref, rtol=1e-12, atol=1e-12)
    # derivatives against central differences
    h = 1e-5
    dV = orc.aziz_values(r, which=1)
    num = (orc.aziz_values(r + h) - orc.aziz_values(r - h)) / (2 * h)
    np.testing.assert_allclose(dV, num, rtol=1e-6, atol=1e-6)
    d2V = orc.aziz_values(r, which=2)
    num2 = (orc.aziz_values(r + h, which=1) - orc.aziz_values(r - h, which=1)) / (2 * h)
    np.testing.assert_allclose(d2V, num2, rtol=1e-6, atol=1e-6)
    assert orc.aziz_values(np.array([0.0]))[0] == 0.0                    # x < EPS
    for year, e in ((1987, 10.948), (1995, 10.956)):
        assert abs(orc.aziz_values(r, year=year).min() + e) < 0.01


def test_aziz_table_and_reference_batch_rule(orc):
    """The reference's own check (tools/benchmarks/potential_benchmark.cpp:103-150,172-241): batched V equals
    scalar V to 1e-9 on its sampleVector inputs, reproduced with the table `direct` lookup."""
    side = np.full(3, synth.C1.side[0])
    max_sep = orc.max_sep(side)
    V, dV, dr = orc.aziz_table(max_sep)
    assert dr == 1.0e-6 * 2.9673 and len(V) == int(max_sep / dr)
    assert V[0] == 0.0 and V[1] > 1e6                                    # hard-core branch
    k = 1_000_000                                                        # r_k accumulated, close to k*dr
    assert abs(V[k] - orc.aziz_values(np.array([k * dr]))[0]) < 1e-7
    n = 4096
    i = np.arange(n)
    t = (i % n) / (n - 1)
    radius = 0.25 + (5.0 - 0.25) * t
    samples = np.stack([radius * np.sin(0.731 * (i + 1) * (d + 1)) for d in range(3)], axis=1)
    samples = np.vstack([samples, [[0.1, 0.2, 1.0], [3.1, 0.0, 1.9], [-4.0, 2.5, -3.0], [5.5, 0.25, 6.75],
                                   [-6.0, -3.0, 7.5], [0.0, 6.5, -6.6]]])
    scalar = np.array([orc.table_V(V, dr, s[None])[0] for s in samples])
    for bs in (1, 2, 3, 7, 16, 64, 257):
        got = np.concatenate([orc.table_V(V, dr, samples[o:o + bs]) for o in range(0, len(samples), bs)])
        err = np.abs(got - scalar)
        assert np.all((err <= 1e-9) | (err / np.maximum(1.0, np.abs(scalar)) <= 1e-9))
    rn = np.linalg.norm(samples, axis=1)
    kk = (rn / dr).astype(int)
    inside = (kk > 0) & (kk < len(V))
    np.testing.assert_array_equal(scalar[inside], V[kk[inside]])
    assert np.all(scalar[~inside] == 0.0)                               # extV = {0,0}


def test_pair_sums_small(orc):
    s = synth.C1
    beads = synth.gen_config(s.N, s.M, s.ndim, s.rho, s.T)
    V, dV, dr = orc.aziz_table(orc.max_sep(s.side))
    dSep = 0.5 * math.sqrt(3) * s.side[2] / 50                          # src/action.cpp:192
    vint, f2, hist = orc.pair_sums(s.side, beads, s.N, V, dV, dr, dSep)
    assert hist.sum(axis=1).max() <= s.N * (s.N - 1) // 2
    # brute-force numpy check of one slice
    t = 7
    pos = beads[t, :s.N]
    v = 0.0
    F = np.zeros((s.N, 3))
    for i in range(s.N):
        for j in range(s.N):
            if i == j:
                continue
            d = pos[i] - pos[j]
            d = d - s.side * np.floor(d / s.side + 0.5)
            r = math.sqrt(d @ d)
            k = int(r / dr)
            if 0 < k < len(V):
                if j > i:
                    v += V[k]
                F[i] += dV[k] / r * d
    assert vint[t] == pytest.approx(v, rel=1e-12)
    assert f2[t] == pytest.approx(np.sum(F * F), rel=1e-11)
    # threads do not change results
    v2, f22, h2 = orc.pair_sums(s.side, beads, s.N, V, dV, dr, dSep, nthreads=3)
    assert np.array_equal(vint, v2) and np.array_equal(f2, f22) and np.array_equal(hist, h2)
    # potentialAction for gsf factors (src/setup.cpp:1238-1248) vs the formula
    VF, GF = [2 / 3, 4 / 3], [0.0, 2 / 9]
    tau, lam = s.tau, synth.LAMBDA_HE4
    U = orc.potential_action(vint, f2, VF, GF, tau, lam)
    ref = sum(VF[t % 2] * tau * vint[t] + (GF[t % 2] * tau**3 * lam * f2[t] if t % 2 else 0.0) for t in range(s.M))
    assert U == pytest.approx(ref, rel=1e-13)
    assert orc.deriv_potential_action_tau(vint[1], f2[1], 1, VF, GF, tau, lam) == pytest.approx(
        VF[1] * vint[1] + 3 * GF[1] * tau * tau * lam * f2[1], rel=1e-14)
    assert orc.deriv_potential_action_lambda(f2[2], 2, GF, tau) == 0.0
    assert orc.deriv_potential_action_lambda(f2[3], 3, GF, tau) == pytest.approx(GF[1] * tau**3 * f2[3])


def test_output_formatting(orc):
    row = orc.format_row([1.0, -2.5e-3, 123456.789], [0.5, 0.5, 0.5], 2)
    assert row == "  2.50000000E-01 -6.25000000E-04  3.08641973E+04"
    assert orc.dvec_to_string([0.5, -1.25, 3.0]) == "(+5.00000000E-01,-1.25000000E+00,+3.00000000E+00)"


def test_energy_estimator_free_particle_closed_form(orc):
    """One free particle hopping by +-d along x every slice: K = NDIM/(2 tau) - d^2/(4 lambda tau^2), V = tail only."""
    M, d, tau, lam = 8, 0.01, 0.05, 6.0
    side = np.array([10.0, 10.0, 10.0])
    beads = np.zeros((M, 1, 3))
    beads[:, 0, 0] = d * (np.arange(M) % 2) - 4.0         # zig-zag: every link, the closing one included, has length d
    zeros = np.zeros(M)
    e = orc.energy(side, beads, 1, zeros, zeros, [1.0, 1.0], [0.0, 0.0], 1, tau, lam, tailV=-2.0, mu=0.3)
    K = 1.5 / tau - d * d / (4.0 * lam * tau * tau)
    Vt = -2.0 / 1000.0
    np.testing.assert_allclose(e, [K, Vt, 0.0, Vt, K + Vt, K + Vt - 0.3, K, Vt, K + Vt], rtol=1e-13, atol=1e-15)
