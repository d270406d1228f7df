"""Golden vectors produced by THE REFERENCE'S OWN CODE (tests/golden/upstream_*.npz, written by
tests/golden/make_upstream_golden.py from the upstream function bodies compiled into oracle/_ref): seeded inputs plus what
the upstream estimators, action, q-generators and Aziz class return for them.  Unlike tests/test_reference_cpu.py these
need neither the upstream tree nor oracle/_ref at test time.  CPU: the oracle reproduces them (bit for bit where the
operation order is the same).  GPU: the CUDA path matches them to 1e-10, integers exactly."""
import math
import os

import numpy as np
import pytest

from parity import assert_parity

HERE = os.path.dirname(os.path.abspath(__file__))
QSETS = {3: [("int", "1 0 0  0 -2 3  5 5 5"), ("float", "0.1 0.25 1.7  -2.2 0.3 0.0"), ("max_int", "2 1 2"), ("max_float", "0.9 0.0 0.0")],
         2: [("int", "1 0 0 1 -1 1"), ("max_int", "3 2"), ("max_float", "1.1 0.0"), ("float", "0.5 -0.25")]}
ACTIONS = {"gsf": ([2 / 3, 4 / 3], [0.0, 2 / 9], 2, 1), "lib": ([1.0, 1.0], [1 / 12, 1 / 12], 2, -1)}


def load(name):
    return np.load(os.path.join(HERE, "golden", name + ".npz"))


@pytest.mark.parametrize("name", ["upstream_3d", "upstream_2d"])
def test_oracle_reproduces_upstream_vectors(orc, name):
    g = load(name)
    N, nd, M = int(g["N"]), int(g["ndim"]), int(g["M"])
    side, beads, q = g["side"], g["beads"], g["q"]
    assert np.array_equal(orc.ssf(side, beads, N, q), g["ssf"])
    assert np.array_equal(orc.isf(beads, N, q), g["isf"])
    for k, (qt, text) in enumerate(QSETS[nd]):
        assert np.array_equal(orc.qvectors(qt, text, side), g[f"qset{k}"])
    dq = 2.0 * math.pi / side[-1]
    for geom in ("line", "sphere"):
        sh = orc.qvectors2(nd, dq, 1.0, geom)
        assert np.array_equal(np.array([len(x) for x in sh]), g[f"q2_{geom}_sizes"]) and np.array_equal(np.vstack(sh), g[f"q2_{geom}"])
    if nd != 3:
        return
    raw, n_in = orc.ssf_cyl(side, beads, N, g["cyl_q"], float(g["cyl_maxR"]))
    assert n_in == int(g["cyl_n1d"]) and np.array_equal(raw / n_in, g["cyl"])
    V, dV, d2V, dr = orc.aziz_table(orc.max_sep(side), second=True)
    assert len(V) == int(g["table_len"]) and dr == float(g["dr"]) and orc.aziz_tail(side[2]) == float(g["tail"])
    idx = g["probe_idx"]
    assert np.array_equal(V[idx], g["probe_V"]) and np.array_equal(dV[idx], g["probe_dV"]) and np.array_equal(d2V[idx], g["probe_d2V"])
    dSep = 0.5 * math.sqrt(3) * side[2] / 50
    vint, f2, hist = orc.pair_sums(side, beads, N, V, dV, dr, dSep)
    tau, lam, window, mu = float(g["tau"]), float(g["lam"]), int(g["window"]), float(g["mu"])
    for a, (VF, GF, period, t2p) in ACTIONS.items():
        assert np.array_equal(hist, g[f"{a}_sephist"]) and np.array_equal(vint, g[f"{a}_vint"]) and np.array_equal(f2, g[f"{a}_f2"])
        f2m = f2.copy()
        for eo in (0, 1):
            if not GF[eo] > 1e-7:
                f2m[eo::2] = 0.0
        np.testing.assert_allclose(orc.potential_action(vint, f2m, VF, GF, tau, lam), float(g[f"{a}_potentialAction"]), rtol=1e-14)
        vir = orc.virial_sums(side, beads, N, window, dV, d2V, dr, t2_parity=t2p, next_links=g["next"])
        eo = np.arange(M) % 2
        c1, c2 = np.array(VF)[eo] * tau, 2.0 * np.array(GF)[eo] * tau ** 3 * lam
        for col, c in ((0, c1), (1, c2), (2, c1), (3, c2)):
            assert_parity(vir[:, col] * c, g[f"{a}_vir"][:, col], f"{a} virial term {col}")
        en = orc.energy(side, beads, N, vint, f2m, VF, GF, period, tau, lam, float(g["tail"]), mu=mu, next_links=g["next"])
        np.testing.assert_allclose(en, g[f"{a}_energy"], rtol=1e-12, atol=1e-12 * np.max(np.abs(g[f"{a}_energy"])))
        ve = orc.virial_energy(side, beads, N, window, vir, vint, f2m, VF, GF, tau, lam, float(g["tail"]), mu=mu, next_links=g["next"])
        np.testing.assert_allclose(ve, g[f"{a}_virial"], rtol=1e-11, atol=1e-11 * np.max(np.abs(g[f"{a}_virial"])))


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["upstream_3d", "upstream_2d"])
def test_cuda_matches_upstream_vectors(orc, name):
    from pimc_b200 import api
    g = load(name)
    N, nd, M = int(g["N"]), int(g["ndim"]), int(g["M"])
    side, beads, q = g["side"], g["beads"], g["q"]
    with api.Context(0, nd) as ctx:
        ctx.set_box(side)
        ctx.set_qvecs(q)
        ssf, isf = ctx.stage(beads, N).ssf_isf()
        assert_parity(ssf[0], g["ssf"], name + ": S(q) vs upstream vectors")
        assert_parity(isf[0], g["isf"], name + ": F(q,tau) vs upstream vectors")
        if nd != 3:
            return
        ctx.set_qvecs(g["cyl_q"])
        cyl, n_in = ctx.ssf_cyl(float(g["cyl_maxR"]))
        assert n_in[0] == int(g["cyl_n1d"])
        assert_parity(cyl[0] / n_in[0], g["cyl"], "cylinder S(q) vs upstream vectors")
        V, dV, d2V, dr = orc.aziz_table(orc.max_sep(side), second=True)      # host-side tables, pinned in the CPU test above
        ctx.set_pair_table(V, dV, dr)
        ctx.set_pair_table_d2(d2V)
        dSep = 0.5 * math.sqrt(3) * side[2] / 50
        vint, f2, hist = ctx.pair_sums(dSep)
        assert np.array_equal(hist[0], g["gsf_sephist"])
        assert_parity(vint[0], g["gsf_vint"], "Vint vs upstream vectors")
        assert_parity(f2[0], g["gsf_f2"], "gradVSquared vs upstream vectors")
        tau, lam, window = float(g["tau"]), float(g["lam"]), int(g["window"])
        delta = orc.virial_delta(side, beads, N, window, next_links=g["next"])
        eo = np.arange(M) % 2
        for a, (VF, GF, period, t2p) in ACTIONS.items():
            vir = ctx.virial_sums(delta, t2_parity=t2p)
            c1, c2 = np.array(VF)[eo] * tau, 2.0 * np.array(GF)[eo] * tau ** 3 * lam
            for col, c in ((0, c1), (1, c2), (2, c1), (3, c2)):
                assert_parity(vir[0][:, col] * c, g[f"{a}_vir"][:, col], f"{a} virial term {col} vs upstream vectors")


@pytest.mark.gpu
def test_cuda_full_c2_against_upstream_cpu_record(orc):
    """The FULL C2 evaluation (N=256, M=170, 64 q; 1.2e11 pair terms on the CPU) as computed once by the upstream CPU code
    (tests/golden/make_upstream_c2_full.py, ~40 core-minutes) against the CUDA path: every q, every tau."""
    path = os.path.join(HERE, "golden", "upstream_c2_full.npz")
    if not os.path.exists(path):
        pytest.skip("tests/golden/upstream_c2_full.npz not generated")
    from pimc_b200 import api, synth
    g = np.load(path)
    s = synth.C2
    beads = synth.gen_config(s.N, s.M, 3, s.rho, s.T, seed=int(g["seed"]), pad=3)
    assert float(np.sum(beads[:, :s.N] * np.arange(1, 4))) == float(g["beads_checksum"]), "synthetic generator changed"
    q = synth.commensurate_q(s.nq, s.side)
    for mode in (1, 0):
        with api.Context(0, 3) as ctx:
            ctx.set_box(s.side)
            ctx.set_qvecs(q)
            ctx.set_rho_mode(mode)
            ssf, isf = ctx.stage(beads, s.N).ssf_isf()
        assert_parity(ssf[0], g["ssf"], f"full C2 S(q) vs the upstream CPU record (rho mode {mode})")
        assert_parity(isf[0], g["isf"], f"full C2 F(q,tau) vs the upstream CPU record (rho mode {mode})")


def test_oracle_full_c2_factorised_against_upstream_cpu_record(orc):
    """CPU: the factorised oracle variant (the checker used at sizes where the direct loop takes hours) against the same
    full-size upstream record."""
    path = os.path.join(HERE, "golden", "upstream_c2_full.npz")
    if not os.path.exists(path):
        pytest.skip("tests/golden/upstream_c2_full.npz not generated")
    from pimc_b200 import synth
    g = np.load(path)
    s = synth.C2
    beads = synth.gen_config(s.N, s.M, 3, s.rho, s.T, seed=int(g["seed"]), pad=3)
    q = synth.commensurate_q(s.nq, s.side)
    assert_parity(orc.isf_factorised(beads, s.N, q), g["isf"], "factorised oracle vs the full-size upstream CPU record")
