"""world_size-2 (and 3) gloo runs of the sharding logic on CPU: q-sharding reproduces the single-process result
bit for bit per shard; configuration sharding + reduce matches the single-process bin to rounding."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, mode, ret):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle
    from pimc_b200 import multi, synth
    orc = oracle.get()
    s = synth.Shape("g", 3, 12, 10, 2.0, 0.02198, 0)
    q = synth.commensurate_q(7, s.side)
    nq = len(q)
    try:
        if mode == "q":
            beads = synth.gen_config(s.N, s.M, 3, s.rho, s.T, seed=3)
            lo, hi = multi.shard_range(nq, world, rank)
            local = np.concatenate([orc.ssf(s.side, beads, s.N, q[lo:hi])[:, None], orc.isf(beads, s.N, q[lo:hi])], axis=1)
            full = multi.gather_q_shards(torch.from_numpy(local), nq).numpy()
            ref = np.concatenate([orc.ssf(s.side, beads, s.N, q)[:, None], orc.isf(beads, s.N, q)], axis=1)
            ret[rank] = bool(np.array_equal(full, ref))
        else:
            B = 5
            batch = synth.gen_batch(s, B)
            lo, hi = multi.shard_range(B, world, rank)
            loc = np.zeros(nq + nq * s.M)
            for b in range(lo, hi):
                loc[:nq] += orc.ssf(s.side, batch[b], s.N, q)
                loc[nq:] += orc.isf(batch[b], s.N, q).reshape(-1)
            tot, n = multi.reduce_bins(torch.from_numpy(loc), hi - lo, dst=0, deterministic=(mode == "cfg_det"))
            if rank == 0:
                ref = np.zeros_like(loc)
                for b in range(B):
                    ref[:nq] += orc.ssf(s.side, batch[b], s.N, q)
                    ref[nq:] += orc.isf(batch[b], s.N, q).reshape(-1)
                ok = n == B and np.allclose(tot.numpy(), ref, rtol=1e-13, atol=1e-12)
                ssf, isf = multi.finalize_bin(tot.numpy(), nq, s.M, n)
                ok = ok and ssf.shape == (nq,) and isf.shape == (nq, s.M)
                ret[rank] = bool(ok)
            else:
                ret[rank] = tot is None
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,mode", [(2, "q"), (3, "q"), (2, "cfg"), (2, "cfg_det")])
def test_sharding_over_gloo(world, mode):
    port = 29600 + (os.getpid() + world * 7 + len(mode)) % 300
    ret = mp.get_context("spawn").Manager().dict()
    mp.spawn(_worker, args=(world, port, mode, ret), nprocs=world, join=True)
    assert all(ret.get(r) for r in range(world)), dict(ret)


def test_shard_range_covers_everything():
    from pimc_b200 import multi
    for n in (1, 7, 64, 256):
        for w in (1, 2, 3, 8):
            spans = [multi.shard_range(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = multi.shard_sizes(n, w)
            assert max(sizes) - min(sizes) <= 1 and sum(sizes) == n
