"""Parity of the pair-potential tables against the REFERENCE'S OWN code: the upstream AzizPotential class
(TabulatedPotential<T>::initLookupTable + direct, AzizPotential parameter sets / valueV / valuedVdr / valued2Vdr2 / V /
gradV / grad2V / tail correction), compiled from the upstream tree into oracle/_ref/librefaziz.so (oracle/Makefile target
`ref`, oracle/ref_aziz_extract.py + oracle/ref_aziz_shim.cpp; built by __graft_entry__.build() wherever /root/reference
exists, shipped with the snapshot).  Held against it: the CPU oracle (bit for bit -- both are compiled with
-ffp-contract=off), the stand-alone host builder pimc_b200/host/aziz.* (through pimcb_host_selftest), and the numpy
builder the bench uses.  CPU only: runs in the `not gpu` suite."""
import ctypes as C
import math
import os
import subprocess

import numpy as np
import pytest

from pimc_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_dp = C.POINTER(C.c_double)


@pytest.fixture(scope="module")
def ref():
    path = os.path.join(ROOT, "oracle", "_ref", "librefaziz.so")
    if not os.path.exists(path):
        pytest.skip(f"{path} not built (needs the upstream tree: make -C oracle ref)")
    lib = C.CDLL(path)
    lib.refaziz_create.restype = C.c_void_p
    lib.refaziz_create.argtypes = [C.c_int, C.c_double, C.c_double]
    lib.refaziz_destroy.argtypes = [C.c_void_p]
    lib.refaziz_table_length.argtypes = [C.c_void_p]
    lib.refaziz_dr.restype = C.c_double
    lib.refaziz_dr.argtypes = [C.c_void_p]
    lib.refaziz_tail.restype = C.c_double
    lib.refaziz_tail.argtypes = [C.c_void_p]
    lib.refaziz_tables.argtypes = [C.c_void_p, _dp, _dp, _dp]
    lib.refaziz_values.argtypes = [C.c_void_p, _dp, C.c_int, _dp, _dp, _dp]
    lib.refaziz_lookup.argtypes = [C.c_void_p, _dp, C.c_int, _dp, _dp, _dp]
    return lib


def ptr(a):
    return a.ctypes.data_as(_dp)


def ref_tables(lib, year, max_sep, rc):
    h = lib.refaziz_create(year, max_sep, rc)
    n = lib.refaziz_table_length(h)
    V, dV, d2V = np.zeros(n), np.zeros(n), np.zeros(n)
    lib.refaziz_tables(h, ptr(V), ptr(dV), ptr(d2V))
    return h, V, dV, d2V, lib.refaziz_dr(h), lib.refaziz_tail(h)


@pytest.mark.parametrize("year", [1979, 1987, 1995])
def test_oracle_tables_bit_identical_to_upstream(ref, orc, year):
    side = synth.C1.side
    max_sep = orc.max_sep(side)
    h, V, dV, d2V, dr, tail = ref_tables(ref, year, max_sep, side[2])
    oV, odV, od2V, odr = orc.aziz_table(max_sep, year=year, second=True)
    assert dr == odr and len(V) == len(oV)
    assert np.array_equal(V, oV) and np.array_equal(dV, odV) and np.array_equal(d2V, od2V)
    assert tail == orc.aziz_tail(side[2], year=year)
    assert orc.aziz_rm(year) * 1.0e-6 == dr
    # the analytic functions away from the table abscissae (hard-core branch, damping switch at x = D, far tail)
    r = np.concatenate([[0.0, 1e-9, 0.02, 0.029], np.linspace(0.5, 12.0, 400), [2.9673 * 1.241314, 3.0 * 1.4826]])
    v, dv, d2v = np.zeros_like(r), np.zeros_like(r), np.zeros_like(r)
    ref.refaziz_values(h, ptr(r), len(r), ptr(v), ptr(dv), ptr(d2v))
    assert np.array_equal(v, orc.aziz_values(r, 0, year)) and np.array_equal(dv, orc.aziz_values(r, 1, year))
    assert np.array_equal(d2v, orc.aziz_values(r, 2, year))
    ref.refaziz_destroy(h)


def test_direct_lookups_match_upstream(ref, orc):
    """AzizPotential::V / gradV / grad2V (sqrt(dot) -> int(r/dr) -> table) on the reference benchmark's kind of
    separation vectors, incl. r = 0 (k <= 0 -> extV[0]) and beyond the table (k >= tableLength -> extV[1])."""
    side = synth.C1.side
    max_sep = orc.max_sep(side)
    h, V, dV, d2V, dr, _ = ref_tables(ref, 1979, max_sep, side[2])
    rng = np.random.default_rng(2)
    sep = rng.uniform(-0.5, 0.5, size=(4000, 3)) * side
    sep[0] = 0.0
    sep[1] = [max_sep, max_sep, 0.0]
    v, g, g2 = np.zeros(len(sep)), np.zeros_like(sep), np.zeros(len(sep))
    ref.refaziz_lookup(h, ptr(np.ascontiguousarray(sep)), len(sep), ptr(v), ptr(g), ptr(g2))
    assert np.array_equal(v, orc.table_V(V, dr, sep))
    assert np.array_equal(g2, orc.table_V(d2V, dr, sep))
    rn = np.sqrt((sep[:, 0] * sep[:, 0] + sep[:, 1] * sep[:, 1]) + sep[:, 2] * sep[:, 2])    # dot(): sequential sum
    with np.errstate(invalid="ignore", divide="ignore"):
        gexp = (orc.table_V(dV, dr, sep) / rn)[:, None] * sep
    assert np.array_equal(g[2:], gexp[2:])
    assert v[0] == 0.0 and v[1] == 0.0 and g2[1] == 0.0
    ref.refaziz_destroy(h)


def test_c2_table_and_numpy_builder_against_upstream(ref, orc):
    """The C2 table (6.6 M entries) and the vectorised numpy builder bench.py uploads (ulp-level agreement: numpy's exp)."""
    side = synth.C2.side
    max_sep = orc.max_sep(side)
    h, V, dV, d2V, dr, _ = ref_tables(ref, 1979, max_sep, side[2])
    oV, odV, od2V, odr = orc.aziz_table(max_sep, second=True)
    assert np.array_equal(V, oV) and np.array_equal(dV, odV) and np.array_equal(d2V, od2V) and dr == odr
    nV, ndV, ndr = synth.aziz_table_numpy(max_sep)
    assert ndr == dr and len(nV) == len(V)
    k = slice(int(1.5 / dr), None)                       # physical range; the hard core is ~1e6 K and steep
    np.testing.assert_allclose(nV[k], V[k], rtol=1e-11, atol=1e-12)
    np.testing.assert_allclose(ndV[k], dV[k], rtol=1e-11, atol=1e-12)
    ref.refaziz_destroy(h)


def test_host_builder_against_upstream(ref):
    """pimc_b200/host/aziz.* (what the stand-alone tools upload) vs the upstream class, through pimcb_host_selftest."""
    host = os.path.join(ROOT, "pimc_b200", "host")
    exe = os.path.join(host, "pimcb_host_selftest3d")
    if not os.path.exists(exe):
        subprocess.run(["make", "-C", host, "NDIM=3"], check=True, stdout=subprocess.DEVNULL)
    out = subprocess.run([exe, "16", repr(0.02198), "int", "1 0 0"], check=True, capture_output=True, text=True).stdout
    rep = dict(l.split("=", 1) for l in out.splitlines() if "=" in l and not l.startswith(("q=", "qraw=", "registered=")))
    side = synth.C1.side
    max_sep = math.sqrt(sum((L / 2.0) ** 2 for L in side))
    h, V, dV, d2V, dr, _ = ref_tables(ref, 1979, max_sep, side[2])
    assert int(rep["tableLength"]) == len(V) and float(rep["dr"]) == dr
    assert float(rep["V1e6"]) == pytest.approx(V[1000000], rel=1e-15)
    assert float(rep["dV1e6"]) == pytest.approx(dV[1000000], rel=1e-15)
    assert float(rep["d2V1e6"]) == pytest.approx(d2V[1000000], rel=1e-14)
    assert float(rep["checksum"]) == pytest.approx(float(np.sum(V[::997] * 1e-3 + dV[::997] * 1e-6)), rel=1e-12)
    assert float(rep["checksum2"]) == pytest.approx(float(np.sum(d2V[::997] * 1e-9)), rel=1e-12)
    ref.refaziz_destroy(h)
