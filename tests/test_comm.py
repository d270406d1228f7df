"""The multi-GPU exchange step inside the C ABI (pimcb_comm_* / pimcb_reduce_bins / pimcb_gather_bins_q): NCCL resolved
with dlopen by the library itself.  One-rank communicator on one GPU; two ranks when the box has two GPUs."""
import os
import sys

import numpy as np
import pytest

from pimc_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_unique_id_needs_no_gpu():
    """ncclGetUniqueId through the dlopen'd library: 128 bytes, different on every call."""
    from pimc_b200 import api
    a, b = api.Context.comm_unique_id(), api.Context.comm_unique_id()
    assert len(a) == 128 and len(b) == 128 and a != b


@pytest.mark.gpu
def test_single_rank_communicator(orc):
    from pimc_b200 import api
    s = synth.Shape("c1", 3, 20, 12, 2.0, 0.02198, 0)
    q = synth.commensurate_q(9, s.side)
    batch = synth.gen_batch(s, 6, first=80)
    with api.Context(0, 3) as ctx:
        ctx.set_box(s.side)
        ctx.set_qvecs(q)
        with pytest.raises(api.PimcbError):
            ctx.reduce_bins(0)                                   # no communicator yet
        ctx.comm_init(1, 0, api.Context.comm_unique_id())
        ctx.stage(batch[:4], s.N).measure()
        ctx.stage(batch[4:], s.N).measure()
        ssf0, isf0, n0 = ctx.read_bins()
        assert ctx.reduce_bins(0) == 6 == n0
        ssf1, isf1, n1 = ctx.read_bins()
        assert n1 == 6 and np.array_equal(ssf0, ssf1) and np.array_equal(isf0, isf1)     # sum over one rank
        g_ssf, g_isf = ctx.gather_bins_q([len(q)])
        assert np.array_equal(g_ssf, ssf1) and np.array_equal(g_isf, isf1)
        with pytest.raises(api.PimcbError):
            ctx.gather_bins_q([len(q) + 1])
        # the reduce that waits for nothing: the count arrives with the next read, and measuring on top of the reduced bin
        # picks it up first
        assert ctx.reduce_bins(0, want_total=False) == 0
        assert ctx.read_bins()[2] == 6
        ctx.reduce_bins(0, want_total=False)
        ctx.stage(batch[:2], s.N).measure()
        ssf2, _, n2 = ctx.read_bins()
        assert n2 == 8
        ctx.reduce_bins(0, want_total=False)
        ctx.reset_bins()
        ctx.stage(batch[:1], s.N).measure()
        assert ctx.read_bins()[2] == 1
        # the pipelined exchange: the snapshot of a bin is reduced while the next bin accumulates
        ctx.reset_bins()
        with pytest.raises(api.PimcbError):
            ctx.reduce_bins_end()                                # nothing in flight
        ctx.stage(batch, s.N).measure()
        ctx.reduce_bins_begin(0)
        with pytest.raises(api.PimcbError):
            ctx.reduce_bins_begin(0)                             # one exchange at a time
        ctx.reset_bins()
        ctx.stage(batch[:2], s.N).measure()                      # the next bin, while the first one travels
        ssf3, isf3, n3 = ctx.reduce_bins_end()
        assert n3 == 6 and np.array_equal(ssf3, ssf1) and np.array_equal(isf3, isf1)
        ssf4, _, n4 = ctx.read_bins()
        assert n4 == 2 and not np.array_equal(ssf4, ssf1)
        ctx.comm_destroy()
    ref = sum(orc.ssf(s.side, b, s.N, q) for b in batch)
    np.testing.assert_allclose(ssf1, ref, rtol=1e-10)


@pytest.mark.gpu
def test_idle_rank_joins_the_reduce_with_a_zero_bin():
    """pimcb_init_bins: a rank whose share of a walker batch is empty (nothing measured) still takes part in
    pimcb_reduce_bins -- with zeros -- instead of failing and leaving its peers inside the collective."""
    from pimc_b200 import api
    s = synth.Shape("c0", 3, 20, 12, 2.0, 0.02198, 0)
    q = synth.commensurate_q(5, s.side)
    with api.Context(0, 3) as ctx:
        ctx.set_box(s.side)
        ctx.set_qvecs(q)
        ctx.comm_init(1, 0, api.Context.comm_unique_id())
        with pytest.raises(api.PimcbError):
            ctx.reduce_bins(0)                                   # no bin laid out yet
        ctx.init_bins(s.M)
        assert ctx.reduce_bins(0) == 0
        ssf, isf, n = ctx.read_bins()
        assert n == 0 and not ssf.any() and not isf.any() and isf.shape == (len(q), s.M)
        ctx.stage(synth.gen_batch(s, 3, first=5), s.N).measure()  # the same layout keeps accumulating afterwards
        ctx.init_bins(s.M)                                       # a second call must not clear what is there
        assert ctx.reduce_bins(0) == 3
        ctx.comm_destroy()


def _rank(rank, world, uid, mode, ret):
    sys.path.insert(0, ROOT)
    from pimc_b200 import api, multi
    s = synth.Shape("c2", 3, 20, 12, 2.0, 0.02198, 0)
    q = synth.commensurate_q(9, s.side)
    batch = synth.gen_batch(s, 6, first=80)
    with api.Context(rank, 3) as ctx:
        ctx.set_box(s.side)
        ctx.comm_init(world, rank, uid)
        if mode == "cfg_pipe":
            ctx.set_qvecs(q)
            lo, hi = multi.shard_range(len(batch), world, rank)
            ctx.stage(batch[lo:hi], s.N).measure()
            ctx.reduce_bins_begin(0)
            ctx.reset_bins()
            ctx.stage(batch[lo:lo + 1], s.N).measure()               # next bin: one configuration per rank
            ssf, isf, n = ctx.reduce_bins_end()
            ctx.reduce_bins_begin(0)
            ssf_b, _, n_b = ctx.reduce_bins_end()
            ret[rank] = (n, 6 if rank == 0 else 3, ssf, isf, n_b, ssf_b)
        elif mode in ("cfg", "cfg_async"):
            ctx.set_qvecs(q)
            lo, hi = multi.shard_range(len(batch), world, rank)
            ctx.stage(batch[lo:hi], s.N).measure()
            n = ctx.reduce_bins(0, want_total=(mode == "cfg"))       # cfg_async: nothing waited for, count with read_bins
            ssf, isf, cnt = ctx.read_bins()
            if mode == "cfg_async":
                n = cnt if rank == 0 else 0
            ret[rank] = (n, cnt, ssf, isf)
        else:
            lo, hi = multi.shard_range(len(q), world, rank)
            ctx.set_qvecs(q[lo:hi])
            ctx.stage(batch, s.N).measure()
            ret[rank] = ctx.gather_bins_q(multi.shard_sizes(len(q), world))
        ctx.comm_destroy()


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["cfg", "cfg_async", "cfg_pipe", "q"])
def test_two_rank_reduce_and_gather(orc, mode):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from pimc_b200 import api
    uid = api.Context.comm_unique_id()
    ret = mp.get_context("spawn").Manager().dict()
    mp.spawn(_rank, args=(2, uid, mode, ret), nprocs=2, join=True)
    s = synth.Shape("c2", 3, 20, 12, 2.0, 0.02198, 0)
    q = synth.commensurate_q(9, s.side)
    batch = synth.gen_batch(s, 6, first=80)
    ssf_ref = sum(orc.ssf(s.side, b, s.N, q) for b in batch)
    isf_ref = sum(orc.isf_factorised(b, s.N, q) for b in batch)
    if mode in ("cfg", "cfg_async", "cfg_pipe"):
        n, cnt, ssf, isf = ret[0][:4]
        assert n == 6 and cnt == 6 and ret[1][0] == 0 and ret[1][1] == 3
        if mode == "cfg_pipe":                                   # the second bin: configurations 0 and 3
            assert ret[0][4] == 2 and ret[1][4] == 0
            np.testing.assert_allclose(ret[0][5], orc.ssf(s.side, batch[0], s.N, q) + orc.ssf(s.side, batch[3], s.N, q), rtol=1e-10)
        np.testing.assert_allclose(ssf, ssf_ref, rtol=1e-10)
        np.testing.assert_allclose(isf, isf_ref, rtol=1e-10, atol=1e-9)
    else:
        for r in range(2):
            np.testing.assert_allclose(ret[r][0], ssf_ref, rtol=1e-10)
            np.testing.assert_allclose(ret[r][1], isf_ref, rtol=1e-10, atol=1e-9)
        assert np.array_equal(ret[0][0], ret[1][0]) and np.array_equal(ret[0][1], ret[1][1])
