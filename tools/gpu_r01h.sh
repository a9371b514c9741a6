#!/bin/bash
TAG=${1:-r01h}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout -s KILL 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_gpu.log
timeout -s KILL 600 python bench.py --steps 10 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open("$OUT/bench.json"))
print({k:d[k] for k in ("value","ms_per_step")}, d["e2e"]["value"], d["e2e_staged"]["value"], d["latency_b1"]["p50_ms"], d["roofline"]["frac"], d["roofline"]["traffic"], d["roofline_small_m"]["frac"])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file $OUT/launches.csv python tools/profile_call.py --graph 0 > $OUT/ncu_launches.log 2>&1; echo "ncu rc=$?"
python tools/summarize_launches.py $OUT/launches.csv > $OUT/launch_summary.txt 2>&1; head -16 $OUT/launch_summary.txt
