#!/bin/bash
TAG=${1:-r01i}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout -s KILL 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_kernels.py -m gpu -x -q -k "gemv or single_observation or ragged or euler or golden" > $OUT/pytest_sel.log 2>&1; echo "pytest rc=$?"; tail -15 $OUT/pytest_sel.log
for G in 1 0; do
VLASER_EXPERT_GEMV=$G timeout -s KILL 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $OUT/bench_gemv$G.json 2> $OUT/bench_gemv$G.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open("$OUT/bench_gemv$G.json"))
print("gemv=$G", {k:d[k] for k in ("value","ms_per_step")}, d["latency_b1"])
PY
done
