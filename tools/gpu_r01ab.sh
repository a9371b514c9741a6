#!/bin/bash
# option prop_stream: proprio pass of the prefill on a second stream beside the LLM layers -- parity and A/B at B=64
OUT=gpurun_out/r01ab; mkdir -p $OUT
VLASER_PROP_STREAM=1 timeout -s KILL 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden or ragged or graph or full_size_2b_parity or dynamic" > $OUT/pytest_sel.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_sel.log
for PS in 1 0 1 0; do
VLASER_PROP_STREAM=$PS timeout -s KILL 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-profile > $OUT/bench_ps$PS.json 2> $OUT/bench_ps$PS.err; echo "bench rc=$?"; tail -2 $OUT/bench_ps$PS.err
python - <<PY
import json
d=json.load(open("$OUT/bench_ps$PS.json"))
print("prop_stream=$PS", {k:d[k] for k in ("value","ms_per_step")}, d["clocks"])
PY
done
