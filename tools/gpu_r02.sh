#!/bin/bash
# One GPU-box visit of round 2: microbenchmarks, the GPU test suite, pair / virial A/B, bench.
TAG=${1:-r02d}
OUT=gpurun_out
mkdir -p $OUT
./tools/micro/pipe_rates > $OUT/${TAG}_pipe_rates.txt 2>&1; tail -20 $OUT/${TAG}_pipe_rates.txt
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > $OUT/${TAG}_pytest_gpu.log; cat $OUT/${TAG}_pytest_gpu.log
echo "== tile + packed tables (default)"; timeout 300 python tools/pair_probe.py 2>&1 | tee $OUT/${TAG}_probe_tile_codec.txt
echo "== tile, verbatim tables"; PIMCB_TABLE_CODEC=0 timeout 300 python tools/pair_probe.py 2>&1 | tee $OUT/${TAG}_probe_tile_raw.txt
echo "== ring kernels (round 1)"; PIMCB_PAIR_TILE=0 PIMCB_VIRIAL_TILE=0 timeout 300 python tools/pair_probe.py 2>&1 | tee $OUT/${TAG}_probe_ring.txt
timeout 600 python bench.py --steps 20 --warmup 5 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; tail -3 $OUT/${TAG}_bench.err; python - <<PY
import json
d = json.load(open("$OUT/${TAG}_bench.json"))
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches")})
print("e2e", d["e2e"]); print("latency", d["latency"])
r = d["roofline"]; print("roofline", {k: r[k] for k in ("achieved", "peak", "frac", "frac_vs_nominal", "avg_launch_ms", "launches_timed", "share_of_step")}, r["corr_kernel"], r.get("upstream_gpu_ab"))
print("pair", d["pair_sums"]); print("cpu", d["cpu_baseline"])
PY
