#!/usr/bin/env python
"""Times pimcb_virial_sums and pimcb_pair_sums (64 C2 configurations, gsf action) with the library PIMCB_LIB_PATH points to."""
import math
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pimc_b200 import api, synth  # noqa: E402

s = synth.C2
B = 64
uniq = synth.gen_batch(s, 8, first=500)
beads = np.stack([uniq[b % 8] for b in range(B)])
max_sep = math.sqrt(sum((L / 2.0) ** 2 for L in s.side))
V, dV, dr = synth.aziz_table_numpy(max_sep)
d2V = np.gradient(dV, dr)
with api.Context(0, 3) as ctx:
    ctx.set_box(s.side)
    ctx.set_pair_table(V, dV, dr)
    ctx.set_pair_table_d2(d2V)
    ctx.stage(beads, s.N)
    delta = 0.01 * beads
    ref = None
    out = {}
    for name, call in (("virial", lambda: ctx.virial_sums(delta, t2_parity=1)),
                       ("pair", lambda: ctx.pair_sums(0.5 * math.sqrt(3.0) * s.side[2] / 50.0, f2_parity=1))):
        call()
        ctx.set_profiling(True)
        ctx.kernel_times(reset=True)
        for _ in range(4):
            res = call()
        ms, n = ctx.kernel_times(reset=True)[name]
        ctx.set_profiling(False)
        out[name] = ms / n
        if name == "virial":
            out["checksum"] = float(np.sum(res))
    print(os.path.basename(os.environ.get("PIMCB_LIB_PATH", "default")), {k: round(v, 6) if k != "checksum" else v for k, v in out.items()})
