"""GPU parity: CUDA path (through the C ABI) vs the CPU oracle on identical seeded inputs."""
import math

import os

import numpy as np
import pytest

from parity import assert_parity
from pimc_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def api():
    from pimc_b200 import api as _api
    return _api


def make_ctx(api, shape, q, periodic=None):
    ctx = api.Context(0, shape.ndim)
    ctx.set_box(shape.side, periodic)
    ctx.set_qvecs(q)
    return ctx


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_c1_full_direct_oracle(api, orc, nthreads, mode):
    """C1 (N=16, M=124, 64 q): every q, every tau against the reference's O(Nq M^2 N^2) loop."""
    s = synth.C1
    beads = synth.gen_config(s.N, s.M, s.ndim, s.rho, s.T)
    q = orc.qvectors("int", synth.int_wavevector_text(s.nq, 3), s.side)
    with make_ctx(api, s, q) as ctx:
        assert ctx.num_commensurate() == s.nq
        ctx.set_rho_mode(mode)
        ssf, isf = ctx.stage(beads, s.N).ssf_isf()
    assert_parity(ssf[0], orc.ssf(s.side, beads, s.N, q, nthreads=nthreads), "C1 ssf")
    assert_parity(isf[0], orc.isf(beads, s.N, q, nthreads=nthreads), "C1 isf")


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_c2_shape(api, orc, nthreads, mode):
    """C2 (N=256, M=170, 64 q): S(q) for all q against the min-image CPU loop; F(q,tau) for all q against the
    factorised CPU variant and for 1 q against the direct reference loop (1.9e9 terms)."""
    s = synth.C2
    beads = synth.gen_config(s.N, s.M, s.ndim, s.rho, s.T, seed=synth.BASE_SEED + 1)
    q = synth.commensurate_q(s.nq, s.side)
    with make_ctx(api, s, q) as ctx:
        ctx.set_rho_mode(mode)
        ssf, isf = ctx.stage(beads, s.N).ssf_isf()
        ssf_only = ctx.ssf()
        isf_only = ctx.isf()
    assert np.array_equal(ssf, ssf_only) and np.array_equal(isf, isf_only)
    assert_parity(ssf[0], orc.ssf(s.side, beads, s.N, q, nthreads=nthreads), "C2 ssf")
    assert_parity(isf[0], orc.isf_factorised(beads, s.N, q), "C2 isf (factorised oracle)")
    k = 37
    assert_parity(isf[0, k:k + 1], orc.isf(beads, s.N, q[k:k + 1], nthreads=nthreads), "C2 isf (direct oracle, 1 q)")


def test_c3_2d_full_grid(api, orc, nthreads):
    """C3: NDIM=2, N=128, M=250, max_int "8 8" -> 289 q incl. q=0, odometer order."""
    s = synth.C3
    beads = synth.gen_config(s.N, s.M, s.ndim, s.rho, s.T, seed=synth.BASE_SEED + 2)
    q = orc.qvectors("max_int", "8 8", s.side)
    assert len(q) == 289
    for mode in (0, 1, 2):
        with make_ctx(api, s, q) as ctx:
            ctx.set_rho_mode(mode)
            ssf, isf = ctx.stage(beads, s.N).ssf_isf()
        assert_parity(ssf[0], orc.ssf(s.side, beads, s.N, q, nthreads=nthreads), f"C3 ssf mode {mode}")
        assert_parity(isf[0], orc.isf_factorised(beads, s.N, q), f"C3 isf mode {mode}")
    zero = np.flatnonzero(np.all(q == 0, axis=1))[0]
    np.testing.assert_allclose(ssf[0, zero], s.M * s.N, rtol=1e-14)


def test_c4_shape_and_q_shards(api, orc, nthreads):
    """C4 (N=1024, M=320, 256 q): all q against the factorised CPU variant, S(q) of 2 q against the min-image CPU loop,
    a handful of F(q,tau) elements against the direct reference loop, and the 8-way q-sharding of BASELINE config 4
    (32 q per shard) reproducing the unsharded columns exactly."""
    from pimc_b200 import multi
    s = synth.C4
    beads = synth.gen_config(s.N, s.M, s.ndim, s.rho, s.T, seed=synth.BASE_SEED + 4)
    q = synth.commensurate_q(s.nq, s.side)
    with make_ctx(api, s, q) as ctx:
        ssf, isf = ctx.stage(beads, s.N).ssf_isf()
    assert_parity(isf[0], orc.isf_factorised(beads, s.N, q), "C4 isf (factorised oracle)")
    assert_parity(ssf[0], isf[0][:, 0], "C4 S(q) = F(q,0)")
    pick = [3, 200]
    assert_parity(ssf[0, pick], orc.ssf(s.side, beads, s.N, q[pick], nthreads=nthreads), "C4 ssf (min-image oracle, 2 q)")
    e0 = 77 * s.M + 150                                  # q 77, tau 150..153: 4 elements of the O(M N^2) loop each
    ref = orc.isf_range(beads, s.N, q, e0, e0 + 4, nthreads=nthreads)[e0:e0 + 4]
    assert_parity(isf[0].reshape(-1)[e0:e0 + 4], ref, "C4 isf (direct oracle, 4 elements)")
    np.testing.assert_allclose(isf[0][:, 1:], isf[0][:, :0:-1], rtol=0, atol=0)      # F(tau) = F(M - tau): mirrored, bit-equal
    for rank in (0, 5, 7):
        lo, hi = multi.shard_range(s.nq, 8, rank)
        with make_ctx(api, s, q[lo:hi]) as ctx:
            s_loc, f_loc = ctx.stage(beads, s.N).ssf_isf()
        assert_parity(f_loc[0], isf[0, lo:hi], f"C4 q-shard {rank}")
        assert_parity(s_loc[0], ssf[0, lo:hi], f"C4 q-shard {rank} ssf")


def test_c5_walker_batch_bin(api, orc):
    """C5 per GPU: a batch of C2 walkers through pimcb_measure (quad-summed correlation + in-kernel bin accumulation)
    equals the sum of the per-configuration results; ragged batch sizes; two configurations against the oracle."""
    s = synth.C2
    q = synth.commensurate_q(s.nq, s.side)
    batch = synth.gen_batch(s, 11, first=50)
    with make_ctx(api, s, q) as ctx:
        ssf, isf = ctx.stage(batch, s.N).ssf_isf()
        for nb in (11, 8, 5, 1):
            ctx.reset_bins()
            ctx.stage(batch[:nb], s.N)
            ctx.measure()
            ctx.measure()                                 # twice: the kernel re-arms its own completion counters
            bs, bi, n = ctx.read_bins()
            assert n == 2 * nb
            assert_parity(bs, 2 * ssf[:nb].sum(axis=0), f"bin ssf B={nb}")
            assert_parity(bi, 2 * isf[:nb].sum(axis=0), f"bin isf B={nb}")
    for b in (0, 10):
        assert_parity(isf[b], orc.isf_factorised(batch[b], s.N, q), f"C5 isf config {b}")


def test_non_commensurate_q_uses_direct_min_image(api, orc, nthreads):
    """`float` wave-vectors: S(q) must follow the CPU min-image pair sum, F(q,tau) raw positions."""
    s = synth.Shape("mix", 3, 64, 40, 2.0, 0.02198, 0)
    beads = synth.gen_config(s.N, s.M, 3, s.rho, s.T, seed=11)
    q = np.vstack([synth.float_q(9, 3), synth.commensurate_q(5, s.side), synth.float_q(2, 3, seed=9)])
    with make_ctx(api, s, q) as ctx:
        assert ctx.num_commensurate() == 5
        ssf, isf = ctx.stage(beads, s.N).ssf_isf()
    assert_parity(ssf[0], orc.ssf(s.side, beads, s.N, q, nthreads=nthreads), "mixed ssf")
    assert_parity(isf[0], orc.isf(beads, s.N, q, nthreads=nthreads), "mixed isf")


@pytest.mark.parametrize("ndim,N,M,pad", [(1, 7, 6, 0), (2, 33, 10, 5), (3, 1, 4, 2), (3, 2, 2, 0), (3, 257, 8, 1)])
def test_ragged_sizes(api, orc, ndim, N, M, pad):
    """Odd particle counts, N=1, N_ext == N, N not a multiple of the CTA width."""
    rho = {1: 0.2, 2: 0.0432, 3: 0.02198}[ndim]
    s = synth.Shape("r", ndim, N, M, 2.0, rho, 0)
    beads = synth.gen_config(N, M, ndim, rho, 2.0, seed=5, pad=pad)
    if pad:
        beads[:, N:, :] = 7777.0          # padding columns must never contribute
    q = np.vstack([synth.commensurate_q(6, s.side, include_zero=True), synth.float_q(3, ndim)])
    for mode in (0, 1, 2):
        with make_ctx(api, s, q) as ctx:
            ctx.set_rho_mode(mode)
            ssf, isf = ctx.stage(beads, N).ssf_isf()
        assert_parity(ssf[0], orc.ssf(s.side, beads, N, q), f"ragged ssf {ndim}D N={N}")
        assert_parity(isf[0], orc.isf(beads, N, q, nthreads=4), f"ragged isf {ndim}D N={N}")


@pytest.mark.parametrize("ndim,nvec", [
    (3, [[1, 0, 0], [0, 0, 1]]),                                   # nmax = (1,0,1): columns with b = 0 only
    (3, [[0, 3, 0], [0, -3, 0], [2, 3, -1], [-2, -3, 1], [2, -3, 1]]),   # anisotropic, nmax = (2,3,1)
    (3, [[4, 1, 0], [0, 0, 4], [1, 1, 1], [-1, 1, 1], [1, -1, 1], [1, 1, -1], [0, 0, 0]]),   # nmax 4: run-time-bound kernel
    (3, [[1, 2, 3], [1, 2, 3], [-1, -2, -3]]),                      # a repeated q opens a second group
    (2, [[0, 5], [5, 0], [-5, 0], [3, -4], [0, 0]]),
    (1, [[1], [-1], [3], [0]]),
    # |n_y| >= 9 / |n_x| >= 9 in 3-D: beyond the fixed stride of the DMMA plan's column map -- no DMMA plan is built
    # (the planner used to index that map out of range, e.g. --wavevector "0 12 0" or max_int "10 10 10")
    (3, [[0, 12, 0], [1, 9, 2], [-2, 10, 1], [0, 0, 1]]),
    (3, [[10, 10, 10], [-10, 9, 3], [12, 0, 0], [3, -11, 16]]),
    (2, [[0, 17], [9, 12], [-12, 3]]),
])
def test_lattice_qsets(api, orc, ndim, nvec):
    """Hand-picked lattice q-sets that exercise every branch of the lattice plans (zero components, anisotropic
    nmax, repeated q, q = 0) in the DMMA (1) and CUDA-core (2) lattice kernels against the generic kernel and oracle."""
    rho = {1: 0.2, 2: 0.0432, 3: 0.02198}[ndim]
    N, M = 40, 6
    s = synth.Shape("l", ndim, N, M, 2.0, rho, 0)
    beads = synth.gen_config(N, M, ndim, rho, 2.0, seed=17)
    q = (2.0 * math.pi / s.side) * np.array(nvec, dtype=float)
    ref_s, ref_f = orc.ssf(s.side, beads, N, q), orc.isf(beads, N, q, nthreads=4)
    for mode in (0, 1, 2):
        with make_ctx(api, s, q) as ctx:
            assert ctx.num_commensurate() == len(q)
            ctx.set_rho_mode(mode)
            ssf, isf = ctx.stage(beads, N).ssf_isf()
        assert_parity(ssf[0], ref_s, f"qset ssf mode {mode}")
        assert_parity(isf[0], ref_f, f"qset isf mode {mode}")


@pytest.mark.parametrize("ndim,seed", [(1, 1), (2, 2), (2, 3), (3, 4), (3, 5), (3, 6), (3, 7), (3, 8)])
def test_random_lattice_qsets_all_rho_kernels_agree(api, orc, ndim, seed):
    """Fuzz of the lattice planner (groups, coinciding factors, zero planes, tile shapes up to 16 x 1 / 8 x 2 / 4 x 4,
    the host-resolved unfold table): random integer q-sets with repeats and q = 0; DMMA (1), CUDA-core lattice (2) and
    generic (0) kernels must agree with each other and with the oracle."""
    rng = np.random.default_rng(1000 + seed)
    rho = {1: 0.2, 2: 0.0432, 3: 0.02198}[ndim]
    N, M = 37, 4
    s = synth.Shape("fz", ndim, N, M, 2.0, rho, 0)
    beads = synth.gen_config(N, M, ndim, rho, 2.0, seed=seed)
    nmax = rng.integers(1, {1: 12, 2: 9, 3: 5}[ndim], size=ndim)
    if seed % 2 == 0:
        nmax[rng.integers(0, ndim)] = 0                    # a dimension that never appears
    nq = int(rng.integers(3, {1: 14, 2: 60, 3: 120}[ndim]))
    n = np.stack([rng.integers(-nmax[d], nmax[d] + 1, size=nq) for d in range(ndim)], axis=1)
    n[0] = 0                                               # q = 0
    n[-1] = n[1]                                           # a repeated q
    q = (2.0 * math.pi / s.side) * n
    res = {}
    for mode in (0, 1, 2):
        with make_ctx(api, s, q) as ctx:
            assert ctx.num_commensurate() == nq
            ctx.set_rho_mode(mode)
            res[mode] = ctx.stage(beads, N).ssf_isf()
    ref_f = orc.isf_factorised(beads, N, q)
    for mode in (0, 1, 2):
        assert_parity(res[mode][1][0], ref_f, f"fuzz isf mode {mode} nmax {nmax.tolist()} nq {nq}")
        assert_parity(res[mode][0][0], ref_f[:, 0], f"fuzz ssf mode {mode}")


@pytest.mark.parametrize("M", [1, 3, 7, 171])
def test_odd_slice_counts(api, orc, M):
    """The reference forces an even number of time slices (src/setup.cpp:1000-1008); the ABI does not: odd M through
    both correlation kernels, the per-configuration and the bin path."""
    N = 6
    s = synth.Shape("odd", 3, N, M, 2.0, 0.02198, 0)
    batch = np.stack([synth.gen_config(N, M, 3, s.rho, 2.0, seed=71 + b) for b in range(3)])
    q = synth.commensurate_q(7, s.side, include_zero=True)
    ref = np.array([orc.isf(batch[b], N, q, nthreads=2) for b in range(3)])
    for mode in (0, 1, 2):
        with make_ctx(api, s, q) as ctx:
            ctx.set_corr_mode(mode)
            ssf, isf = ctx.stage(batch, N).ssf_isf()
            ctx.reset_bins()
            ctx.measure()
            bs, bi, n = ctx.read_bins()
        for b in range(3):
            assert_parity(isf[b], ref[b], f"odd M={M} isf corr mode {mode}")
            assert_parity(ssf[b], ref[b][:, 0], f"odd M={M} ssf corr mode {mode}")
        assert n == 3
        assert_parity(bi, ref.sum(axis=0), f"odd M={M} bin corr mode {mode}")


@pytest.mark.parametrize("M", [2, 4, 6, 14, 62, 126, 128, 130, 254, 258, 382, 386, 510, 512, 640])
def test_tau_correlation_kernels(api, orc, M):
    """The tau-correlation kernels (1 = DMMA, 2 = DMMA with persistent CTAs and the next pair prefetched, both up to
    M = 510 then falling back; 0 = CUDA cores) over time-slice counts that straddle every 64-tau accumulator tile boundary
    of the DMMA formulation, two configurations per batch."""
    N = 5
    s = synth.Shape("corr", 3, N, M, 2.0, 0.02198, 0)
    batch = np.stack([synth.gen_config(N, M, 3, s.rho, 2.0, seed=31 + b) for b in range(2)])
    q = np.vstack([synth.commensurate_q(5, s.side, include_zero=True), synth.float_q(2, 3)])
    out = {}
    for mode in (0, 1, 2):
        with make_ctx(api, s, q) as ctx:
            ctx.set_corr_mode(mode)
            out[mode] = ctx.stage(batch, N).ssf_isf()
    assert np.array_equal(out[1][1], out[2][1]) and np.array_equal(out[1][0], out[2][0]), "modes 1 and 2: same arithmetic"
    for b in range(2):
        ref_f = orc.isf_factorised(batch[b], N, q)
        ref_s = orc.ssf(s.side, batch[b], N, q)
        for mode in (0, 1, 2):
            assert_parity(out[mode][1][b], ref_f, f"isf corr mode {mode} M={M}")
            assert_parity(out[mode][0][b], ref_s, f"ssf corr mode {mode} M={M}")
    if M <= 62:
        assert_parity(out[1][1][0], orc.isf(batch[0], N, q, nthreads=4), f"isf direct M={M}")


def test_tau_correlation_persistent_ctas_walk_many_items(api, orc):
    """corr mode 2 with more work items than resident CTAs (1776 quads x q on 148 x <= 7 CTAs): every CTA loops, the
    prefetched pair of item n + 1 must not leak into item n, the double-buffered parking area of the bin path must hold.
    Per-configuration results and the bin are bit-identical to mode 1; a sample of configurations against the oracle."""
    N, M = 6, 76                                        # M >= 64 + 11: the persistent kernel's two-image staging table applies
    s = synth.Shape("pipe", 3, N, M, 2.0, 0.02198, 0)
    uniq = synth.gen_batch(s, 37, first=500)
    B = 1481                                            # not a multiple of 4: the last quad has dead warps
    batch = np.ascontiguousarray(uniq[np.arange(B) % len(uniq)])
    q = synth.commensurate_q(5, s.side, include_zero=True)
    res = {}
    for mode in (1, 2):
        with make_ctx(api, s, q) as ctx:
            ctx.set_corr_mode(mode)
            ssf, isf = ctx.stage(batch, N).ssf_isf()
            ctx.reset_bins()
            ctx.measure()
            ctx.measure()                               # accumulates on top of the persistent rows
            res[mode] = (ssf, isf) + ctx.read_bins()
    for a, b in zip(res[1], res[2]):
        assert np.array_equal(a, b)
    assert res[2][4] == 2 * B
    for b in (0, 36, 1480):
        assert_parity(res[2][1][b], orc.isf_factorised(batch[b], N, q), f"persistent corr, configuration {b}")
    assert_parity(res[2][3], 2.0 * res[2][1].sum(axis=0), "bin = sum of the per-configuration rows, twice")


@pytest.mark.parametrize("split", [1, 2, 3, 5])
def test_rho_split_slices(api, orc, split, monkeypatch):
    """Persistent-warp rho kernel with `split` warps sharing every slice (partial tiles parked in global scratch, last
    warp of a slice reduces in fixed order): same numbers for every split, launch after launch (self re-arming counters)."""
    N, M = 150, 12                                     # 5 particle blocks per slice, the last one ragged
    s = synth.Shape("split", 3, N, M, 2.0, 0.02198, 0)
    batch = synth.gen_batch(s, 3)
    q = synth.commensurate_q(20, s.side)
    monkeypatch.setenv("PIMCB_RHO_SPLIT", str(split))
    with make_ctx(api, s, q) as ctx:
        first = ctx.stage(batch, N).ssf_isf()
        again = ctx.stage(batch, N).ssf_isf()
    assert np.array_equal(first[0], again[0]) and np.array_equal(first[1], again[1]), "deterministic across launches"
    for b in range(3):
        assert_parity(first[0][b], orc.ssf(s.side, batch[b], N, q), f"split {split} ssf")
        assert_parity(first[1][b], orc.isf_factorised(batch[b], N, q), f"split {split} isf")


def test_batch_bins_and_slots(api, orc):
    """A walker batch: per-configuration outputs, device-resident bin accumulation, slot rotation."""
    s = synth.Shape("b", 3, 32, 16, 2.0, 0.02198, 0)
    q = synth.commensurate_q(10, s.side)
    batch = synth.gen_batch(s, 5)
    with make_ctx(api, s, q) as ctx:
        ssf, isf = ctx.stage(batch, s.N).ssf_isf()
        for b in range(5):
            assert_parity(ssf[b], orc.ssf(s.side, batch[b], s.N, q), f"batch ssf {b}")
            assert_parity(isf[b], orc.isf(batch[b], s.N, q), f"batch isf {b}")
        ctx.reset_bins()
        ctx.measure()
        ctx.stage(batch[:2], s.N, slot=2)
        ctx.select_slot(2)
        ctx.measure()
        bs, bi, n = ctx.read_bins()
        assert n == 7
        assert_parity(bs, ssf.sum(axis=0) + ssf[:2].sum(axis=0), "bins ssf")
        assert_parity(bi, isf.sum(axis=0) + isf[:2].sum(axis=0), "bins isf")
        ctx.reset_bins()
        ctx.measure()
        assert ctx.read_bins()[2] == 2
        assert ctx.launch_count() > 0


def test_pinned_source_takes_device_transpose_path(api, orc):
    s = synth.Shape("p", 3, 48, 12, 2.0, 0.02198, 0)
    q = synth.commensurate_q(8, s.side)
    batch = synth.gen_batch(s, 3, pad=5)
    pin = api.PinnedArray(batch.shape)
    pin.array[...] = batch
    with make_ctx(api, s, q) as ctx:
        a = ctx.stage(batch, s.N).ssf_isf()
        b = ctx.stage(pin.array, s.N).ssf_isf()
        # asynchronous staging of page-locked sources: two batches enqueued back to back, results read afterwards
        pin2 = api.PinnedArray(batch.shape)
        pin2.array[...] = batch[::-1]
        ctx.stage_async(pin.array, s.N)
        c1 = ctx.ssf_isf()
        ctx.stage_async(pin2.array, s.N)
        ctx.stage_async(pin.array, s.N)
        ctx.stage_wait()
        c2 = ctx.ssf_isf()
        ctx.reset_bins()
        ctx.stage_async(pin2.array, s.N)
        ctx.measure()
        bs, bi, n = ctx.read_bins()
        # fused stage + evaluate + read-back of one configuration, page-locked and pageable source
        f1 = ctx.ssf_isf_beads(pin.array[1], s.N)
        f2 = ctx.ssf_isf_beads(batch[1].copy(), s.N)
        for f in (f1, f2):
            assert np.array_equal(f[0][0], a[0][1]) and np.array_equal(f[1][0], a[1][1])
    pin.free()
    pin2.free()
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    assert np.array_equal(a[0], c1[0]) and np.array_equal(a[1], c1[1])
    assert np.array_equal(a[0], c2[0]) and np.array_equal(a[1], c2[1])
    assert n == 3
    assert_parity(bi, a[1].sum(axis=0), "bin after async staging")


@pytest.mark.parametrize("shape", [synth.C1, synth.C2])
def test_pair_sums(api, orc, nthreads, shape):
    """Vint[M] (+ sepHist, bit-exact) and gradVSquared[M] against LocalAction::V / gradVSquared restated on the CPU,
    with the host-built Aziz table; total potentialAction for the gsf factors."""
    s = shape
    beads = synth.gen_config(s.N, s.M, s.ndim, s.rho, s.T, seed=synth.BASE_SEED + 3)
    V, dV, dr = orc.aziz_table(orc.max_sep(s.side))
    dSep = 0.5 * math.sqrt(3) * s.side[2] / 50
    cv, cf, ch = orc.pair_sums(s.side, beads, s.N, V, dV, dr, dSep, nthreads=nthreads)
    with api.Context(0, 3) as ctx:
        ctx.set_box(s.side)
        ctx.set_pair_table(V, dV, dr)
        ctx.stage(beads, s.N)
        gv, gf, gh = ctx.pair_sums(dSep)
        gv2, _, gh2 = ctx.pair_sums(dSep, want_f2=False)
        gv3, gf3, _ = ctx.pair_sums(dSep, want_hist=False, f2_parity=1)
    assert np.array_equal(gh[0], ch), "sepHist must be bit-exact"
    assert np.array_equal(gh2[0], ch)
    assert_parity(gv[0], cv, "Vint")
    assert_parity(gv2[0], cv, "Vint (V-only kernel)")
    assert_parity(gf[0], cf, "gradVSquared")
    assert_parity(gv3[0], cv, "Vint (odd-slice f2)")
    assert_parity(gf3[0, 1::2], cf[1::2], "gradVSquared odd slices")
    assert np.all(gf3[0, 0::2] == 0.0)
    VF, GF = [2 / 3, 4 / 3], [0.0, 2 / 9]
    Ug = orc.potential_action(gv3[0], gf3[0], VF, GF, s.tau, synth.LAMBDA_HE4)
    Uc = orc.potential_action(cv, cf, VF, GF, s.tau, synth.LAMBDA_HE4)
    assert abs(Ug - Uc) <= 1e-10 * abs(Uc)


def test_pair_table_index_is_exact_at_bin_boundaries(api, orc):
    """The tile kernel takes k = int(r/dr) and the sepHist bin from a multiplication by the rounded reciprocal plus a
    safety margin, and re-runs the reference's exact operation sequence inside the margin.  With V[k] = k the slice sum
    IS the index: two particles per slice at separations (k0 + delta) dr with delta from far inside the margin to far
    outside it (and the same around the histogram's bin edges); Vint and sepHist must equal the oracle's exactly."""
    side = np.array([40.0, 40.0, 40.0])
    dr = 2.9673e-6
    nk = 3_000_000
    V = np.arange(nk, dtype=np.float64)
    dV = -np.arange(nk, dtype=np.float64)
    dSep = 0.5 * math.sqrt(3) * side[2] / 50
    rng = np.random.default_rng(7)
    deltas = np.array([0.0, 1e-13, -1e-13, 1e-11, -1e-11, 1e-9, -1e-9, 1e-7, -1e-7, 5e-7, -5e-7, 1e-6, -1e-6, 2e-6, -2e-6, 1e-4, -1e-4,
                       0.25, 0.5, 0.999999, 0.9999999999])
    seps = []
    for k0 in rng.integers(700_000, nk - 10, size=60):
        seps += [(k0 + d) * dr for d in deltas]
    for b in range(1, 14):                                   # histogram edges (bins of dSep), well inside the table
        seps += [b * dSep * (1.0 + e) for e in (0.0, 1e-16, -1e-16, 1e-15, -1e-15, 1e-12, -1e-12, 1e-9, -1e-9) if b * dSep < (nk - 2) * dr]
    seps = np.array(seps)
    M = len(seps)
    beads = np.zeros((M, 2, 3))
    beads[:, 0, 0] = -0.3
    beads[:, 1, 0] = -0.3 + seps                             # along x: r = |dx| up to the rounding of the subtraction
    half = M // 2                                            # second half: along a diagonal (all components in the norm)
    beads[half:, 1, 0] = -0.3 + seps[half:] / math.sqrt(3)
    beads[half:, 1, 1] = seps[half:] / math.sqrt(3)
    beads[half:, 1, 2] = -seps[half:] / math.sqrt(3)
    cv, cf, ch = orc.pair_sums(side, beads, 2, V, dV, dr, dSep)
    with api.Context(0, 3) as ctx:
        ctx.set_box(side)
        ctx.set_pair_table(V, dV, dr)
        ctx.stage(beads, 2)
        gv, gf, gh = ctx.pair_sums(dSep)
        gv2, _, gh2 = ctx.pair_sums(dSep, want_f2=False)
    assert np.array_equal(gv[0], cv) and np.array_equal(gv2[0], cv), np.flatnonzero(gv[0] != cv)[:10]
    assert np.array_equal(gh[0], ch) and np.array_equal(gh2[0], ch)
    assert_parity(gf[0], cf, "gradVSquared with V[k] = k tables")


@pytest.mark.parametrize("ndim,N,M,pad,per", [(3, 1, 3, 0, None), (3, 2, 5, 1, None), (3, 33, 7, 2, None), (3, 95, 6, 0, None),
                                              (3, 130, 5, 3, (1, 1, 0)), (2, 128, 9, 1, None), (2, 47, 4, 0, (1, 0)),
                                              (1, 40, 6, 2, None), (3, 300, 3, 1, None), (3, 700, 2, 0, None)])
def test_pair_sums_ragged_shapes(api, orc, nthreads, ndim, N, M, pad, per):
    """Group counts from 1 to 22 (partial last group, odd and even numbers of groups: full tiles, shared half tiles,
    several rounds of partner slots, several slices per CTA), 1-D / 2-D, slab periodicity, gsf parity and all slices."""
    rho = {1: 0.2, 2: 0.0432, 3: 0.02198}[ndim]
    s = synth.Shape("rg", ndim, N, M, 2.0, rho, 0)
    beads = synth.gen_config(N, M, ndim, rho, 2.0, seed=500 + N, pad=pad)
    maxsep = math.sqrt(sum((L / 2) ** 2 for L in s.side)) * (2.0 if per is not None else 1.0)
    V, dV, dr = orc.aziz_table(maxsep)
    dSep = 0.5 * math.sqrt(ndim) * s.side[-1] / 50
    periodic = np.array(per, dtype=np.uint32) if per is not None else None
    cv, cf, ch = orc.pair_sums(s.side, beads, N, V, dV, dr, dSep, periodic=periodic, nthreads=nthreads)
    with api.Context(0, ndim) as ctx:
        ctx.set_box(s.side, periodic)
        ctx.set_pair_table(V, dV, dr)
        ctx.stage(beads, N)
        gv, gf, gh = ctx.pair_sums(dSep)
        gv1, gf1, gh1 = ctx.pair_sums(dSep, f2_parity=1)
        gv0, _, _ = ctx.pair_sums(dSep, want_f2=False, want_hist=False)
    assert np.array_equal(gh[0], ch) and np.array_equal(gh1[0], ch)
    for g in (gv, gv1, gv0):
        assert_parity(g[0], cv, f"Vint N={N}")
    assert_parity(gf[0], cf, f"gradVSquared N={N}")
    assert_parity(gf1[0, 1::2], cf[1::2], "odd slices")
    assert np.all(gf1[0, 0::2] == 0.0)


@pytest.mark.parametrize("M", [1, 5, 6])
@pytest.mark.parametrize("parity", [0, 1])
def test_pair_sums_parity_split_over_a_batch(api, orc, nthreads, M, parity):
    """A gsf-type call (gradVSquared on the slices of one parity) is two launches over disjoint slices: the force kernel on
    the selected parity, the V-only kernel on the rest.  Batches of several configurations with odd and even slice counts:
    the slice selection must follow (t mod 2) inside every configuration, every output row must be written exactly once,
    and the histogram of every slice must be there whichever launch produced it."""
    N, B = 70, 3
    s = synth.Shape("split", 3, N, M, 2.0, 0.02198, 0)
    batch = np.stack([synth.gen_config(N, M, 3, s.rho, 2.0, seed=640 + b) for b in range(B)])
    V, dV, dr = orc.aziz_table(orc.max_sep(s.side))
    dSep = 0.5 * math.sqrt(3) * s.side[2] / 50
    with api.Context(0, 3) as ctx:
        ctx.set_box(s.side)
        ctx.set_pair_table(V, dV, dr)
        ctx.stage(batch, N)
        gv, gf, gh = ctx.pair_sums(dSep, f2_parity=parity)
    for b in range(B):
        cv, cf, ch = orc.pair_sums(s.side, batch[b], N, V, dV, dr, dSep, nthreads=nthreads)
        assert np.array_equal(gh[b], ch), f"sepHist, configuration {b}"
        assert_parity(gv[b], cv, f"Vint, configuration {b}")
        if len(cf[parity::2]):
            assert_parity(gf[b, parity::2], cf[parity::2], f"gradVSquared on parity {parity}, configuration {b}")
        assert np.all(gf[b, 1 - parity::2] == 0.0)


def test_pair_table_edges(api, orc):
    """Separations below dr (k <= 0 -> extV[0]) and beyond the table (k >= len -> extV[1])."""
    side = np.array([30.0, 30.0, 30.0])
    M, N = 2, 4
    beads = np.zeros((M, N, 3))
    beads[:, 0] = [0.0, 0.0, 0.0]
    beads[:, 1] = [1e-7, 0.0, 0.0]           # r < dr
    beads[:, 2] = [14.0, 14.0, 14.0]         # r ~ 24 A, beyond a short table
    beads[:, 3] = [3.0, 0.1, -0.2]
    V, dV, dr = orc.aziz_table(12.0)
    ext = (5.0, -7.0)
    import numpy as _np
    cv = _np.zeros(M)
    with api.Context(0, 3) as ctx:
        ctx.set_box(side)
        ctx.set_pair_table(V, dV, dr, extV=ext, extdVdr=(0.0, 0.0))
        ctx.stage(beads, N)
        gv, _, gh = ctx.pair_sums(0.5 * math.sqrt(3) * 30.0 / 50, want_f2=False)
    # oracle with the same extremal values
    from oracle import oracle as _o
    lib = _o.get().lib
    import ctypes as C
    sidec, per = _np.ascontiguousarray(side), _np.ones(3, dtype=_np.uint32)
    extc = _np.array(ext)
    hist = _np.zeros((M, 50), dtype=_np.int32)
    dp = C.POINTER(C.c_double)
    rc = lib.orc_pair_sums(3, sidec.ctypes.data_as(dp), per.ctypes.data_as(C.POINTER(C.c_uint)),
                           beads.ctypes.data_as(dp), M, N, N, V.ctypes.data_as(dp), dV.ctypes.data_as(dp), len(V), dr,
                           extc.ctypes.data_as(dp), extc.ctypes.data_as(dp), 0.5 * math.sqrt(3) * 30.0 / 50,
                           cv.ctypes.data_as(dp), None, hist.ctypes.data_as(C.POINTER(C.c_int)), 1)
    assert rc == 0
    assert_parity(gv[0], cv, "edge Vint")
    assert np.array_equal(gh[0], hist)


def test_error_paths(api):
    with pytest.raises(api.PimcbError):
        api.Context(0, 4)
    with api.Context(0, 3) as ctx:
        with pytest.raises(api.PimcbError):
            ctx.set_qvecs(np.zeros((2, 3)))          # box first
        ctx.set_box([5.0, 5.0, 5.0])
        ctx.set_qvecs(np.ones((2, 3)))
        with pytest.raises(api.PimcbError):
            ctx.shape = (1, 4, 2)
            ctx.ssf()                                 # nothing staged
        with pytest.raises(api.PimcbError):
            ctx.stage(np.zeros((4, 2, 3)), 3)         # N > N_ext
        ctx.stage(np.zeros((4, 2, 3)), 2)
        with pytest.raises(api.PimcbError):
            ctx.pair_sums(1.0)                        # no table


@pytest.mark.parametrize("N,pad", [(45, 3), (64, 0), (150, 5)])
def test_graph_replay_leaves_the_transposed_beads_for_the_pair_kernels(api, orc, nthreads, N, pad):
    """In the captured single-walker call the rho_q kernel reads the page-locked beads array itself and writes the
    transposed copy on the way (no transpose kernel).  Whatever runs next on the same ctx -- pimcb_pair_sums here, what
    LocalActionB200 does after an estimator's accumulate() -- must find exactly the staged layout: ragged last particle
    block, padded source rows (N_ext > N), zero padding up to Npad."""
    M = 10
    s = synth.Shape("direct", 3, N, M, 2.0, 0.02198, 0)
    q = synth.commensurate_q(12, s.side)
    cfgs = np.stack([synth.gen_config(N, M, 3, s.rho, 2.0, seed=900 + k, pad=pad) for k in range(4)])
    V, dV, dr = orc.aziz_table(orc.max_sep(s.side))
    dSep = 0.5 * math.sqrt(3.0) * s.side[2] / 50.0
    pa = api.PinnedArray(cfgs.shape[1:])
    try:
        with make_ctx(api, s, q) as ctx, make_ctx(api, s, q) as plain:
            ctx.set_pair_table(V, dV, dr)
            plain.set_pair_table(V, dV, dr)
            for k in range(4):
                pa.array[...] = cfgs[k]
                n0 = ctx.launch_count()
                a = ctx.ssf_isf_beads(pa.array, N)
                if k >= 2 and os.environ.get("PIMCB_RHO_DIRECT") != "0":
                    assert ctx.launch_count() - n0 == 2                      # replay: rho_q + tau-correlation
                b = plain.stage(cfgs[k], N).ssf_isf()
                assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
                gv, gf, gh = ctx.pair_sums(dSep)                             # on the copy the rho kernel left behind
                pv, pf, ph = plain.pair_sums(dSep)
                assert np.array_equal(gv, pv) and np.array_equal(gf, pf) and np.array_equal(gh, ph), f"call {k}"
            cv, cf, ch = orc.pair_sums(s.side, cfgs[3], N, V, dV, dr, dSep, nthreads=nthreads)
            assert_parity(gv[0], cv, "Vint after a graph replay")
            assert np.array_equal(gh[0], ch)
    finally:
        pa.free()


def test_fused_single_walker_call_graph_replay(api, orc, nthreads):
    """pimcb_ssf_isf_beads with a page-locked source: call 1 runs the ordinary path, call 2 captures the CUDA graph,
    later calls replay it.  Every call must return exactly what the ordinary path returns for the CURRENT contents of the
    source buffer; a new q-set, a new shape or a pageable source fall back and re-capture."""
    s = synth.C1
    q = synth.commensurate_q(24, s.side)
    cfgs = synth.gen_batch(s, 6, first=300)
    pa = api.PinnedArray(cfgs.shape[1:])
    try:
        with make_ctx(api, s, q) as ctx, make_ctx(api, s, q) as plain:
            got = []
            for k in range(6):
                pa.array[...] = cfgs[k]
                n0 = ctx.launch_count()
                got.append(ctx.ssf_isf_beads(pa.array, s.N))
                if k >= 2:
                    # rho_q (reading the page-locked beads array itself, leaving the transposed copy behind) and the
                    # tau-correlation per replay; PIMCB_RHO_DIRECT=0 puts the transpose kernel back in front
                    assert ctx.launch_count() - n0 == (3 if os.environ.get("PIMCB_RHO_DIRECT") == "0" else 2)
                ref_ssf, ref_isf = plain.stage(cfgs[k], s.N).ssf_isf()
                assert np.array_equal(got[-1][0], ref_ssf) and np.array_equal(got[-1][1], ref_isf), f"call {k}"
            assert_parity(got[5][0][0], orc.ssf(s.side, cfgs[5], s.N, q, nthreads=nthreads), "graph replay S(q)")
            assert_parity(got[5][1][0], orc.isf(cfgs[5], s.N, q, nthreads=nthreads), "graph replay F(q,tau)")
            # the elastic-scattering epilogue finds the replayed results in the cache
            es = ctx.elastic()
            np.testing.assert_allclose(es[0], 2.0 / s.M * got[5][1][0][:, :s.M // 2 + 1].sum(axis=1), rtol=1e-12)
            # new q-set: dropped, primed, captured again
            q2 = synth.commensurate_q(9, s.side)
            ctx.set_qvecs(q2)
            plain.set_qvecs(q2)
            for k in range(4):
                pa.array[...] = cfgs[k]
                a = ctx.ssf_isf_beads(pa.array, s.N)
                b = plain.stage(cfgs[k], s.N).ssf_isf()
                assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
            # pageable source and a different shape in between
            a = ctx.ssf_isf_beads(cfgs[1].copy(), s.N)
            b = plain.stage(cfgs[1], s.N).ssf_isf()
            assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
            half = np.ascontiguousarray(cfgs[2][: s.M // 2])
            a = ctx.ssf_isf_beads(half, s.N)
            b = plain.stage(half, s.N).ssf_isf()
            assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
            for k in range(3):
                pa.array[...] = cfgs[k + 3]
                a = ctx.ssf_isf_beads(pa.array, s.N)
                b = plain.stage(cfgs[k + 3], s.N).ssf_isf()
                assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    finally:
        pa.free()
