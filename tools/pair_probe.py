#!/usr/bin/env python
"""How much of the pair / virial kernels' time is the table footprint?  Same C2 batch, same kernels, tables subsampled
by 1, 2, 4, 16 (dr multiplied, entries kept verbatim): 53 MB per table -> 26 / 13 / 3.3 MB.  Results are NOT parity
results (the coarse tables are not the reference's), only timings."""
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pimc_b200 import api, synth  # noqa: E402

s = synth.SHAPES[os.environ.get("WORKLOAD", "C2")]
B = int(os.environ.get("BATCH", "64"))
uniq = synth.gen_batch(s, 8, first=0)
pa = api.PinnedArray((B,) + uniq.shape[1:])
for b in range(B):
    pa.array[b] = uniq[b % len(uniq)]
max_sep = math.sqrt(sum((L / 2.0) ** 2 for L in s.side))
V, dV, d2V, dr = synth.aziz_table_numpy(max_sep, second=True)
dSep = 0.5 * math.sqrt(3.0) * s.side[2] / 50.0
with api.Context(0, s.ndim) as ctx:
    ctx.set_box(s.side)
    ctx.stage(pa.array, s.N)
    for sub in [int(x) for x in os.environ.get("SUBS", "1,2,16").split(",")]:
        ctx.set_pair_table(np.ascontiguousarray(V[::sub]), np.ascontiguousarray(dV[::sub]), dr * sub)
        ctx.set_pair_table_d2(np.ascontiguousarray(d2V[::sub]))
        out = {}
        for what in ("pair_gsf", "pair_gsf_nohist", "pair_vonly", "pair_vonly_nohist", "pair_allf", "virial"):
            ctx.set_profiling(True)
            for it in range(4):
                if it == 1:
                    ctx.kernel_times(reset=True)
                if what.startswith("pair_gsf"):
                    ctx.pair_sums(dSep, want_f2=True, want_hist="nohist" not in what, f2_parity=1)
                elif what.startswith("pair_vonly"):
                    ctx.pair_sums(dSep, want_f2=False, want_hist="nohist" not in what)
                elif what == "pair_allf":
                    ctx.pair_sums(dSep, want_f2=True, want_hist=True, f2_parity=-1)
                else:
                    ctx.virial_sums(0.01 * pa.array, t2_parity=1)
            kt = ctx.kernel_times(reset=True)
            ms, n = kt["pair"] if what != "virial" else kt["virial"]
            out[what] = ms / max(1, n)
            ctx.set_profiling(False)
        print(f"table/{sub:<2d} ({8 * len(V[::sub]) / 1e6:6.1f} MB per table): " + "  ".join(f"{k} {v:6.3f}" for k, v in out.items()) + "  (ms)", flush=True)
pa.free()
