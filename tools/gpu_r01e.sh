#!/bin/bash
# Staging-kernel parity + A/B of the per-kernel expert schedule switches. Usage: bash tools/gpu_r01e.sh <tag>
TAG=${1:-r01e}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout -s KILL 600 python -m pytest tests/test_gpu_staging.py -m gpu -x -q > $OUT/pytest_staging.log 2>&1; echo "pytest rc=$?"; tail -15 $OUT/pytest_staging.log
run() {
  NAME=$1; shift
  env "$@" timeout -s KILL 600 python bench.py --steps 10 --warmup 3 > $OUT/bench_$NAME.json 2> $OUT/bench_$NAME.err; echo "bench $NAME rc=$?"
  python - <<PY
import json
d=json.load(open("$OUT/bench_$NAME.json"))
print("$NAME", {k:d[k] for k in ("value","ms_per_step")}, d["e2e"]["value"], d["latency_b1"]["p50_ms"])
PY
}
run mega0 VLASER_EXPERT_MEGA=0
run mega0_fuse VLASER_EXPERT_MEGA=0 VLASER_FUSE_NORM=1
run mega0_pdl2 VLASER_EXPERT_MEGA=0 VLASER_PDL=2
run mega0_fuse_pdl2 VLASER_EXPERT_MEGA=0 VLASER_FUSE_NORM=1 VLASER_PDL=2
run mega1_pdl2 VLASER_EXPERT_MEGA=1 VLASER_PDL=2
