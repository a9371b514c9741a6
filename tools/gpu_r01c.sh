#!/bin/bash
# Expert-stream iteration: parity tests, per-phase timing under A/B flags, bench line. Usage: bash tools/gpu_r01c.sh <tag> [flags]
TAG=${1:-r01c}
FLAGS=${2:-0,16,32,14,30}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout -s KILL 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > $OUT/pytest_parity.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_parity.log
timeout -s KILL 300 python tools/expert_phases.py --flags $FLAGS > $OUT/expert_phases.txt 2>&1; grep -E "====|phases,|mean" $OUT/expert_phases.txt | grep -v "proprio\|last" | head -80
timeout -s KILL 600 python bench.py --steps 10 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open("$OUT/bench.json"))
print({k:d[k] for k in ("value","ms_per_step")}, d["e2e"]["value"], d["latency_b1"], d.get("roofline_expert",{}).get("avg_launch_us"))
PY
