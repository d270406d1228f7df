#!/usr/bin/env python
"""Per-kernel device times of the single-walker path (B = 1) and the bare H2D / D2H times of its buffers."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pimc_b200 import api, synth  # noqa: E402

for name in ("C1", "C2", "C4"):
    s = synth.SHAPES[name]
    q = synth.commensurate_q(s.nq, s.side)
    beads = synth.gen_config(s.N, s.M, s.ndim, s.rho, s.T)
    pa = api.PinnedArray(beads.shape)
    pa.array[...] = beads
    with api.Context(0, s.ndim) as ctx:
        ctx.set_box(s.side)
        ctx.set_qvecs(q)
        for _ in range(3):
            ctx.stage(pa.array, s.N).ssf_isf()
        ctx.set_profiling(True)
        ctx.kernel_times(reset=True)
        for _ in range(20):
            ctx.stage(pa.array, s.N).ssf_isf()
        kt = ctx.kernel_times(reset=True)
        ctx.set_profiling(False)
    d = torch.empty(beads.size, dtype=torch.float64, device="cuda")
    h = torch.from_numpy(pa.array.reshape(-1))
    out_d = torch.empty(len(q) * (1 + s.M), dtype=torch.float64, device="cuda")
    out_h = torch.empty(len(q) * (1 + s.M), dtype=torch.float64).pin_memory()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    for _ in range(3):
        d.copy_(h, non_blocking=True); out_h.copy_(out_d, non_blocking=True)
    torch.cuda.synchronize()
    e0.record(); d.copy_(h, non_blocking=True); e1.record(); out_h.copy_(out_d, non_blocking=True); e2.record()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(200):
        torch.cuda.synchronize()
    sync_us = (time.perf_counter() - t0) / 200 * 1e6
    print(name, {k: round(1e3 * v[0] / max(1, v[1]), 2) for k, v in kt.items() if v[1]}, "us;  H2D %.1f us (%d KB), D2H %.1f us, idle sync %.1f us"
          % (e0.elapsed_time(e1) * 1e3, beads.nbytes // 1024, e1.elapsed_time(e2) * 1e3, sync_us))
    pa.free()
