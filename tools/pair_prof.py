#!/usr/bin/env python
"""Launches for an ncu capture of the pair / virial kernels: C2 batch of 64, V-only pass, gsf pass, virial pass (twice each)."""
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
    ctx.set_pair_table(V, dV, dr)
    ctx.set_pair_table_d2(d2V)
    print(ctx.table_codec_info())
    for _ in range(2):
        ctx.pair_sums(dSep, want_f2=False, want_hist=True)
    for _ in range(2):
        ctx.pair_sums(dSep, want_f2=True, want_hist=True, f2_parity=1)
    if os.environ.get("WITH_VIRIAL", "1") != "0":
        for _ in range(2):
            ctx.virial_sums(0.01 * pa.array, t2_parity=1)
pa.free()
