#!/bin/bash
# A/B: threshold of the CTA-pair GEMM (expert gate/up at M = 256 has 70 pair tiles)
TAG=${1:-r01z}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for PM in 74 64 74 64; do
VLASER_PAIR_MIN_TILES=$PM timeout -s KILL 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-profile > $OUT/bench_pm$PM.json 2> $OUT/bench_pm$PM.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open("$OUT/bench_pm$PM.json"))
print("pair_min_tiles=$PM", {k:d[k] for k in ("value","ms_per_step")}, d["clocks"])
PY
done
