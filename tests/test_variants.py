"""Scattering-family variants and virial slice sums (row f3).

CPU part: known-answer tests of the oracle restatements (getQVectors2, cylinder S(q), elastic scattering, the virial
slice sums and the virial energy estimator).  GPU part: the CUDA path through the C ABI against the oracle.
"""
import math

import numpy as np
import pytest

from parity import assert_parity
from pimc_b200 import synth

LAM = synth.LAMBDA_HE4


# ------------------------------------------------------------------------------------------------------------------
# helpers
# ------------------------------------------------------------------------------------------------------------------
def cylinder_config(N, M, L, R, seed, pad=2):
    """Beads inside a box periodic only along z; radial positions spread across and beyond the cut-off radius."""
    rng = np.random.default_rng(seed)
    side = np.array([2.2 * R, 2.2 * R, L])
    base = np.column_stack([rng.uniform(-R, R, N), rng.uniform(-R, R, N), rng.uniform(-0.5 * L, 0.5 * L, N)])
    walk = np.cumsum(rng.normal(0.0, 0.12, size=(M, N, 3)), axis=0)
    walk -= walk.mean(axis=0, keepdims=True)
    pos = base[None] + walk
    pos[..., 2] -= L * np.floor(pos[..., 2] / L + 0.5)
    beads = np.zeros((M, N + pad, 3))
    beads[:, :N] = pos
    beads[:, N:] = 0.01          # padding columns sit INSIDE the radius: they must not be counted
    return beads, side, np.array([0, 0, 1], dtype=np.uint32)


def permuted_links(M, N, Next, seed):
    """Closed world lines with a random particle permutation across the slice M-1 -> 0 boundary."""
    rng = np.random.default_rng(seed)
    nxt = np.full((M, Next, 2), -1, dtype=np.int32)
    for s in range(M):
        nxt[s, :N, 0] = (s + 1) % M
        nxt[s, :N, 1] = np.arange(N)
    nxt[M - 1, :N, 1] = rng.permutation(N)
    return nxt


# ------------------------------------------------------------------------------------------------------------------
# CPU: oracle known answers
# ------------------------------------------------------------------------------------------------------------------
def test_qvectors2_line_and_sphere(orc):
    L = 20.0
    dq = 2.0 * math.pi / L
    line = orc.qvectors2(3, dq, 4.0, "line")
    assert len(line) == int(4.0 / dq + 1e-7) + 1 and all(len(s) == 1 for s in line)
    assert np.all(line[0] == 0.0)
    for k, s in enumerate(line[1:], start=1):
        assert s[0, 0] == 0.0 and s[0, 1] == 0.0
        np.testing.assert_allclose(s[0, 2], k * dq, rtol=1e-14)
    sph = orc.qvectors2(3, dq, 1.0, "sphere")
    assert len(sph[0]) == 1 and len(sph[1]) == len(sph[2]) > 24
    for k, s in enumerate(sph[1:], start=1):
        np.testing.assert_allclose(np.linalg.norm(s, axis=1), k * dq, rtol=1e-13)      # every vector on its shell
        assert np.all(s >= -1e-15)                                                      # positive octant
        assert np.array_equal(s[0], [0.0, 0.0, s[0, 2]])                                # theta = 0 first
    two = orc.qvectors2(2, dq, 1.0, "sphere")                                           # no angular set below 3-D
    assert all(len(s) == 1 for s in two) and two[1][0, 0] == 0.0
    with pytest.raises(ValueError):
        orc.qvectors2(3, dq, 1.0, "plane")


def test_oracle_cylinder_ssf_known_answers(orc):
    beads, side, per = cylinder_config(12, 6, 9.0, 3.0, seed=1)
    N = 12
    q = np.array([[0, 0, k * 2 * math.pi / 9.0] for k in range(5)])
    # radius larger than the box: every bead counts, and for these q the sum is |rho_q|^2 summed over slices
    full, n_in = orc.ssf_cyl(side, beads, N, q, 100.0, periodic=per)
    assert n_in == N
    ph = np.einsum("qd,tnd->qtn", q, beads[:, :N])
    rho2 = np.abs(np.exp(1j * ph).sum(axis=2)) ** 2
    np.testing.assert_allclose(full, rho2.sum(axis=1), rtol=1e-11, atol=1e-9)
    # finite radius: same closed form restricted to the beads inside
    R = 2.0
    inside = (beads[:, :N, 0] ** 2 + beads[:, :N, 1] ** 2) < R * R
    cut, n_in = orc.ssf_cyl(side, beads, N, q, R, periodic=per)
    assert n_in == int(inside[0].sum()) and 0 < n_in < N
    rho2 = np.abs((np.exp(1j * ph) * inside[None]).sum(axis=2)) ** 2
    np.testing.assert_allclose(cut, rho2.sum(axis=1), rtol=1e-11, atol=1e-9)
    assert cut[0] == np.sum(inside.sum(axis=1) ** 2)                                    # q = 0: n_in(t)^2 summed


def test_oracle_elastic_is_half_range_sum_of_isf(orc, nthreads):
    s = synth.Shape("e", 3, 9, 12, 2.0, 0.02198, 0)
    beads = synth.gen_config(s.N, s.M, 3, s.rho, s.T, seed=11)
    q = synth.commensurate_q(6, s.side)
    isf = orc.isf(beads, s.N, q, nthreads=nthreads)                  # isf/N, all tau
    es = orc.elastic(beads, s.N, q, nthreads=nthreads)
    np.testing.assert_allclose(es, 2.0 / s.M * isf[:, :s.M // 2 + 1].sum(axis=1), rtol=1e-12)


def test_oracle_virial_two_particles_closed_form(orc):
    """Two beads per slice: sum gV.r = r dV/dr and, because gV is parallel to the separation, T gV = d2V gV."""
    side = np.array([30.0, 30.0, 30.0])
    V, dV, d2V, dr = orc.aziz_table(orc.max_sep(side), second=True)
    M = 4
    beads = np.zeros((M, 3, 3))
    seps = np.array([[2.7, 0.3, -0.4], [3.1, 1.0, 0.2], [0.5, -3.3, 0.9], [2.0, 2.0, 2.0]])
    beads[:, 0] = np.array([1.0, -2.0, 0.5])
    beads[:, 1] = beads[:, 0] - seps
    out = orc.virial_sums(side, beads, 2, 1, dV, d2V, dr)
    r = np.linalg.norm(seps, axis=1)
    k = (r / dr).astype(int)
    np.testing.assert_allclose(out[:, 0], dV[k] * r, rtol=1e-12)
    np.testing.assert_allclose(out[:, 1], d2V[k] * dV[k] * r, rtol=1e-11)
    # window 1: the centroid is the bead itself, so delta = 0
    np.testing.assert_allclose(out[:, 2:], 0.0, atol=1e-12)
    # parity selection of the T-matrix terms
    odd = orc.virial_sums(side, beads, 2, 1, dV, d2V, dr, t2_parity=1)
    assert np.all(odd[0::2, 1] == 0.0) and np.array_equal(odd[1::2, 1], out[1::2, 1]) and np.array_equal(odd[:, 0], out[:, 0])


def test_oracle_virial_delta_follows_links(orc):
    s = synth.Shape("v", 3, 7, 8, 2.0, 0.02198, 0)
    beads = synth.gen_config(s.N, s.M, 3, s.rho, s.T, seed=5, pad=2)
    Next = beads.shape[1]
    straight = orc.virial_delta(s.side, beads, s.N, 3)
    ident = permuted_links(s.M, s.N, Next, 0)
    ident[s.M - 1, :s.N, 1] = np.arange(s.N)
    assert np.array_equal(orc.virial_delta(s.side, beads, s.N, 3, next_links=ident), straight)
    perm = permuted_links(s.M, s.N, Next, 3)
    d = orc.virial_delta(s.side, beads, s.N, 3, next_links=perm)
    assert np.array_equal(d[3], straight[3])                   # a window of 3 around slice 3 never crosses the boundary
    assert not np.array_equal(d[0], straight[0])
    assert np.all(d[:, s.N:] == 0.0)
    # independent numpy evaluation for straight world lines (window 2: beads t-1, t, t, t+1 in unwrapped coordinates)
    x = beads[:, :s.N]
    fwd = synth.put_in_bc(np.roll(x, -1, axis=0) - x, s.side)
    bwd = synth.put_in_bc(np.roll(x, 1, axis=0) - x, s.side)
    com = (x + x + (x + fwd) + (x + bwd)) / 4.0
    np.testing.assert_allclose(orc.virial_delta(s.side, beads, s.N, 2)[:, :s.N], synth.put_in_bc(x - com, s.side), atol=1e-13)


def test_oracle_virial_energy_thermodynamic_column_matches_energy_estimator(orc, nthreads):
    """E_th of the virial estimator is the thermodynamic energy of the `energy` estimator (same three terms)."""
    s = synth.Shape("ve", 3, 10, 8, 2.0, 0.02198, 0)
    beads = synth.gen_config(s.N, s.M, 3, s.rho, s.T, seed=21, pad=1)
    Next = beads.shape[1]
    links = permuted_links(s.M, s.N, Next, 2)
    V, dV, d2V, dr = orc.aziz_table(orc.max_sep(s.side), second=True)
    dSep = 0.5 * math.sqrt(3) * s.side[2] / 50
    vint, f2, _ = orc.pair_sums(s.side, beads, s.N, V, dV, dr, dSep)
    VF, GF = [2.0 / 3.0, 4.0 / 3.0], [0.0, 2.0 / 9.0]                                   # gsf, src/setup.cpp:1240-1246
    tail = orc.aziz_tail(s.side[2])
    vir = orc.virial_sums(s.side, beads, s.N, 5, dV, d2V, dr, t2_parity=1, next_links=links, nthreads=nthreads)
    ve = orc.virial_energy(s.side, beads, s.N, 5, vir, vint, f2, VF, GF, s.tau, LAM, tail, next_links=links)
    en = orc.energy(s.side, beads, s.N, vint, f2, VF, GF, 2, s.tau, LAM, tail, next_links=links)
    cols = {k: v for k, v in zip(orc.VIRIAL_COLUMNS, ve)}
    np.testing.assert_allclose(cols["E_th"], en[4], rtol=1e-12)
    np.testing.assert_allclose(cols["V_op"], en[3], rtol=1e-12)                          # even slices, period 2
    # the misspelt "cVCov2" key of upstream lands on column 0 (K_op); without the quirk it is a column of its own
    clean = orc.virial_energy(s.side, beads, s.N, 5, vir, vint, f2, VF, GF, s.tau, LAM, tail, next_links=links, quirk=False)
    assert cols["CvCov2"] == 0.0 and clean[15] != 0.0
    np.testing.assert_allclose(ve[0] - clean[0], clean[15], rtol=1e-12)
    np.testing.assert_allclose(clean[0] + clean[2], clean[4], rtol=1e-12)                # K_op + V_op = E
    np.testing.assert_allclose(clean[1] + clean[3], clean[4], rtol=1e-12)                # K_cv + V_cv = E
    np.testing.assert_allclose(cols["Ecv*Beta"], clean[4] * s.M * s.tau, rtol=1e-13)


def test_symmetric_half_ring_index_logic():
    """The every-pair-once force kernels (pair_sym_kernel / virial_sym_kernel) walk the half ring k = 1..N/2 with 256
    threads, PPT particles per thread and two ring steps per group: mirrored here index for index -- every unordered pair
    exactly once, and all partners of one ring step distinct (the read-add-write into the shared-memory accumulators needs
    no atomics)."""
    for N in (1, 2, 3, 16, 37, 255, 256, 257, 300, 511, 512, 700, 1024):
        PPT = 1 if N <= 256 else (2 if N <= 512 else 4)
        khalf, even = N // 2, N % 2 == 0
        seen = {}
        for kk0 in range(1, khalf + 1, 2):
            for w in range(2):
                targets = []
                for tid in range(256):
                    for p in range(PPT):
                        i, kk = tid + 256 * p, kk0 + w
                        if not (i < N and kk <= khalf and not (even and kk == khalf and i >= khalf)):
                            continue
                        j = i + kk - (N if i + kk >= N else 0)
                        targets.append(j)
                        key = (min(i, j), max(i, j))
                        seen[key] = seen.get(key, 0) + 1
                assert len(targets) == len(set(targets)), (N, kk0, w)
        assert len(seen) == N * (N - 1) // 2 and all(v == 1 for v in seen.values()), N


# ------------------------------------------------------------------------------------------------------------------
# GPU: CUDA path vs oracle
# ------------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def api():
    from pimc_b200 import api as _api
    return _api


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["C1", "2d", "mixed"])
def test_elastic_scattering_vs_oracle(api, orc, nthreads, name):
    if name == "C1":
        s, q = synth.C1, synth.commensurate_q(16, synth.C1.side)
    elif name == "2d":
        s = synth.Shape("e2", 2, 33, 21, 1.0, 0.0432, 0)              # odd M
        q = orc.qvectors("max_int", "2 2", s.side)
    else:
        s = synth.Shape("em", 3, 20, 16, 2.0, 0.02198, 0)
        q = np.vstack([synth.commensurate_q(5, s.side), synth.float_q(3, 3)])
    beads = synth.gen_batch(s, 3, first=40)
    with api.Context(0, s.ndim) as ctx:
        ctx.set_box(s.side)
        ctx.set_qvecs(q)
        es = ctx.stage(beads, s.N).elastic()
        _, isf = ctx.ssf_isf()
        launches = ctx.launch_count()
        es2 = ctx.elastic()                                           # cached pass: one more launch only
        assert ctx.launch_count() == launches + 1
    assert np.array_equal(es, es2)
    for b in range(3):
        assert_parity(es[b], orc.elastic(beads[b], s.N, q, nthreads=nthreads), f"{name}: elastic scattering, config {b}")
    np.testing.assert_allclose(es, 2.0 / s.M * isf[:, :, :s.M // 2 + 1].sum(axis=2), rtol=1e-12)


@pytest.mark.gpu
def test_cylinder_ssf_vs_oracle(api, orc):
    """Cylinder S(q): box periodic along z only, the estimator's own "line" q-set (commensurate) plus
    non-commensurate vectors through the masked pair loop; batch of 2 with different occupation of the radius."""
    N, M, L, R = 37, 12, 11.0, 3.0
    cfgs = [cylinder_config(N, M, L, R, seed=k) for k in (3, 4)]
    beads = np.stack([c[0] for c in cfgs])
    side, per = cfgs[0][1], cfgs[0][2]
    shells = orc.qvectors2(3, 2.0 * math.pi / L, 4.0, "line")
    q = np.vstack(shells + [synth.float_q(4, 3, seed=9)])
    maxR = 2.0
    with api.Context(0, 3) as ctx:
        ctx.set_box(side, per)
        ctx.set_qvecs(q)
        assert ctx.num_commensurate() == len(shells)
        out, n_in = ctx.stage(beads, N).ssf_cyl(maxR)
    for b in range(2):
        ref, n_ref = orc.ssf_cyl(side, beads[b], N, q, maxR, periodic=per)
        assert n_in[b] == n_ref and 0 < n_ref < N
        assert_parity(out[b], ref, f"cylinder S(q), config {b}")
        # the estimator's increment: shell sums / num1DParticles (src/estimator.cpp:5455)
        assert_parity(out[b][:len(shells)] / n_in[b], ref[:len(shells)] / n_ref, "estimator increment")


@pytest.mark.gpu
def test_cylinder_ssf_argument_checks(api):
    with api.Context(0, 1) as ctx:
        ctx.set_box([10.0])
        ctx.set_qvecs(np.array([[0.5]]))
        ctx.stage(np.zeros((2, 3, 1)), 3)
        with pytest.raises(api.PimcbError):
            ctx.ssf_cyl(1.0)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["small3d", "2d", "C2"])
def test_virial_sums_vs_oracle(api, orc, nthreads, name):
    if name == "small3d":
        s, window, B = synth.Shape("v3", 3, 37, 10, 2.0, 0.02198, 0), 5, 2
    elif name == "2d":
        s, window, B = synth.Shape("v2", 2, 40, 9, 1.0, 0.0432, 0), 3, 3          # odd M, several configurations
    else:
        s, window, B = synth.C2, 5, 1
    beads = synth.gen_batch(s, B, first=70)
    Next = beads.shape[2]
    V, dV, d2V, dr = orc.aziz_table(orc.max_sep(s.side), second=True)
    links = permuted_links(s.M, s.N, Next, 17)
    delta = np.stack([orc.virial_delta(s.side, beads[b], s.N, window, next_links=links) for b in range(B)])
    with api.Context(0, s.ndim) as ctx:
        ctx.set_box(s.side)
        with pytest.raises(api.PimcbError):
            ctx.stage(beads, s.N).virial_sums(delta)                  # no tables yet
        ctx.set_pair_table(V, dV, dr)
        first_only = ctx.virial_sums(None, t2_parity=-2)              # no d2V table needed without the T-matrix terms
        ctx.set_pair_table_d2(d2V)
        full = ctx.virial_sums(delta, t2_parity=-1)
        odd = ctx.virial_sums(delta, t2_parity=1)
        even = ctx.virial_sums(delta, t2_parity=0)                    # (two launches each: T-matrix kernel + gV-only kernel)
    for b in range(B):
        ref = orc.virial_sums(s.side, beads[b], s.N, window, dV, d2V, dr, t2_parity=-1, next_links=links, nthreads=nthreads)
        for k, what in enumerate(("sum gV.r", "sum (T gV).r", "sum gV.delta", "sum (T gV).delta")):
            assert_parity(full[b, :, k], ref[:, k], f"{name}: {what}, config {b}")
        assert np.array_equal(first_only[b, :, 0], full[b, :, 0]) and np.all(first_only[b, :, 1:] == 0.0)
        assert np.array_equal(odd[b, 1::2], full[b, 1::2])
        assert np.all(odd[b, 0::2, 1] == 0.0) and np.all(odd[b, 0::2, 3] == 0.0)
        assert np.array_equal(odd[b, 0::2, 0], full[b, 0::2, 0]) and np.array_equal(odd[b, 0::2, 2], full[b, 0::2, 2])
        assert np.array_equal(even[b, 0::2], full[b, 0::2])
        assert np.all(even[b, 1::2, 1] == 0.0) and np.all(even[b, 1::2, 3] == 0.0)
        assert np.array_equal(even[b, 1::2, 0], full[b, 1::2, 0]) and np.array_equal(even[b, 1::2, 2], full[b, 1::2, 2])


@pytest.mark.gpu
def test_virial_energy_from_device_sums(api, orc, nthreads):
    """The 19 virial-estimator columns assembled from device sums (pair + virial kernels) against the all-CPU chain."""
    s = synth.Shape("vE", 3, 64, 16, 2.0, 0.02198, 0)
    beads = synth.gen_config(s.N, s.M, 3, s.rho, s.T, seed=91)
    Next = beads.shape[1]
    links = permuted_links(s.M, s.N, Next, 8)
    V, dV, d2V, dr = orc.aziz_table(orc.max_sep(s.side), second=True)
    dSep = 0.5 * math.sqrt(3) * s.side[2] / 50
    VF, GF = [2.0 / 3.0, 4.0 / 3.0], [0.0, 2.0 / 9.0]
    tail = orc.aziz_tail(s.side[2])
    window = 5
    delta = orc.virial_delta(s.side, beads, s.N, window, next_links=links)
    with api.Context(0, 3) as ctx:
        ctx.set_box(s.side)
        ctx.set_pair_table(V, dV, dr)
        ctx.set_pair_table_d2(d2V)
        vint, f2, _ = ctx.stage(beads, s.N).pair_sums(dSep, f2_parity=1)
        vir = ctx.virial_sums(delta, t2_parity=1)
    got = orc.virial_energy(s.side, beads, s.N, window, vir[0], vint[0], f2[0], VF, GF, s.tau, LAM, tail, next_links=links)
    cv, cf, _ = orc.pair_sums(s.side, beads, s.N, V, dV, dr, dSep, nthreads=nthreads)
    cf[0::2] = 0.0
    cvir = orc.virial_sums(s.side, beads, s.N, window, dV, d2V, dr, t2_parity=1, next_links=links, nthreads=nthreads)
    ref = orc.virial_energy(s.side, beads, s.N, window, cvir, cv, cf, VF, GF, s.tau, LAM, tail, next_links=links)
    np.testing.assert_allclose(got, ref, rtol=1e-10, atol=1e-10 * np.max(np.abs(ref)))


@pytest.mark.gpu
@pytest.mark.parametrize("N", [300, 600])
def test_force_kernels_with_several_particles_per_thread(api, orc, nthreads, N):
    """N > 256: the every-pair-once kernels run with 2 / 4 particles per thread (pair sums, sepHist, gradVSquared, virial)."""
    s = synth.Shape("pp", 3, N, 4, 2.0, 0.02198, 0)
    beads = synth.gen_config(s.N, s.M, 3, s.rho, s.T, seed=123, pad=1)
    V, dV, d2V, dr = orc.aziz_table(orc.max_sep(s.side), second=True)
    dSep = 0.5 * math.sqrt(3) * s.side[2] / 50
    window = 2
    delta = orc.virial_delta(s.side, beads, s.N, window)
    with api.Context(0, 3) as ctx:
        ctx.set_box(s.side)
        ctx.set_pair_table(V, dV, dr)
        ctx.set_pair_table_d2(d2V)
        vint, f2, hist = ctx.stage(beads, s.N).pair_sums(dSep)
        vir = ctx.virial_sums(delta, t2_parity=-1)
    cv, cf, ch = orc.pair_sums(s.side, beads, s.N, V, dV, dr, dSep, nthreads=nthreads)
    assert np.array_equal(hist[0], ch)
    assert_parity(vint[0], cv, f"N={N} Vint")
    assert_parity(f2[0], cf, f"N={N} gradVSquared")
    ref = orc.virial_sums(s.side, beads, s.N, window, dV, d2V, dr, t2_parity=-1, nthreads=nthreads)
    for k in range(4):
        assert_parity(vir[0][:, k], ref[:, k], f"N={N} virial term {k}")
