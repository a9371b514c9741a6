#!/bin/bash
# Expert-stream sub-grid iteration. Usage: bash tools/gpu_r01d.sh <tag>
TAG=${1:-r01d}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout -s KILL 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > $OUT/pytest_parity.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_parity.log
timeout -s KILL 300 python tools/expert_phases.py --flags 0 --streams 1,2,4 > $OUT/expert_phases.txt 2>&1; grep -E "####|====|phases,|mean" $OUT/expert_phases.txt | grep -v "last" | head -80
for S in 1 2 4; do
VLASER_EXPERT_STREAMS=$S timeout -s KILL 600 python bench.py --steps 10 --warmup 3 > $OUT/bench_s$S.json 2> $OUT/bench_s$S.err; echo "bench streams=$S rc=$?"
python - <<PY
import json
d=json.load(open("$OUT/bench_s$S.json"))
print({k:d[k] for k in ("value","ms_per_step")}, d["e2e"]["value"], d["latency_b1"], d.get("roofline_expert",{}).get("avg_launch_us"))
PY
done
