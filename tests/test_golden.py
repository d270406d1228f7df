"""Committed golden vectors (tests/golden/*.npz, made by tests/golden/make_golden.py from the oracle):
the oracle must keep reproducing them (CPU), and the CUDA path must match them (GPU)."""
import os

import numpy as np
import pytest

from parity import assert_parity

HERE = os.path.dirname(os.path.abspath(__file__))


def load(name):
    return np.load(os.path.join(HERE, "golden", name + ".npz"))


@pytest.mark.parametrize("name", ["small_3d", "small_2d"])
def test_oracle_reproduces_golden(orc, name):
    g = load(name)
    N = int(g["N"])
    assert np.array_equal(orc.ssf(g["side"], g["beads"], N, g["q"]), g["ssf"])
    assert np.array_equal(orc.isf(g["beads"], N, g["q"]), g["isf"])
    if name == "small_3d":
        V, dV, dr = orc.aziz_table(orc.max_sep(g["side"]))
        assert len(V) == int(g["table_len"]) and dr == float(g["dr"])
        assert np.array_equal(V[g["table_probe_idx"]], g["table_probe_V"])
        assert np.array_equal(dV[g["table_probe_idx"]], g["table_probe_dV"])
        vint, f2, hist = orc.pair_sums(g["side"], g["beads"], N, V, dV, dr, float(g["dSep"]))
        assert np.array_equal(vint, g["vint"]) and np.array_equal(f2, g["f2"]) and np.array_equal(hist, g["hist"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["small_3d", "small_2d"])
def test_cuda_matches_golden(orc, name):
    from pimc_b200 import api
    g = load(name)
    N, nd = int(g["N"]), int(g["ndim"])
    with api.Context(0, nd) as ctx:
        ctx.set_box(g["side"])
        ctx.set_qvecs(g["q"])
        ssf, isf = ctx.stage(g["beads"], N).ssf_isf()
        assert_parity(ssf[0], g["ssf"], name + " ssf")
        assert_parity(isf[0], g["isf"], name + " isf")
        if name == "small_3d":
            V, dV, dr = orc.aziz_table(orc.max_sep(g["side"]))    # table construction is host-side by design
            ctx.set_pair_table(V, dV, dr)
            vint, f2, hist = ctx.pair_sums(float(g["dSep"]))
            assert_parity(vint[0], g["vint"], "golden Vint")
            assert_parity(f2[0], g["f2"], "golden gradVSquared")
            assert np.array_equal(hist[0], g["hist"])
