#!/bin/bash
# A/B of the L2 prefetch riding on the norm launches of the denoise steps (B=64), plus the golden parity tests
TAG=${1:-r01l}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout -s KILL 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden or ragged or graph" > $OUT/pytest_sel.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/pytest_sel.log
for PF in 1 0 1 0; do
VLASER_L2_PREFETCH=$PF timeout -s KILL 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-profile > $OUT/bench_pf$PF.json 2> $OUT/bench_pf$PF.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open("$OUT/bench_pf$PF.json"))
print("l2_prefetch=$PF", {k:d[k] for k in ("value","ms_per_step")}, d["e2e"]["value"], d["clocks"])
PY
done
