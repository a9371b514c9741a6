#!/bin/bash
# closing pass of the round: every GPU test, smoke, bench (both arms) with the committed code
TAG=${1:-r01final}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout -s KILL 1500 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_gpu.log
timeout -s KILL 300 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $OUT/smoke.log
timeout -s KILL 600 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"
timeout -s KILL 300 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err; echo "ref rc=$?"
python - <<PY
import json
d=json.load(open("$OUT/bench.json"))
print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}, d["e2e"]["value"], d["latency_b1"], d["clocks"], d["roofline"]["frac"], d["cpu_baseline"]["value"])
r=json.load(open("$OUT/bench_reference.json")); print("reference arm", r["value"], r["unit"], r["cpu_baseline"]["cores"])
PY
