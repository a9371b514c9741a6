#!/bin/bash
TAG=${1:-r01v}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout -s KILL 1500 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_gpu.log
timeout -s KILL 600 python tools/b1_latency.py > $OUT/b1.txt 2> $OUT/b1.err; echo "b1 rc=$?"; grep "expert_gemv=1 stop" $OUT/b1.txt
timeout -s KILL 600 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open("$OUT/bench.json"))
print({k:d[k] for k in ("value","ms_per_step")}, d["e2e"]["value"], d["latency_b1"], d["clocks"])
PY
