#!/bin/bash
# certification of the final commit of the round: every GPU test, smoke, bench (both arms)
OUT=gpurun_out/r02s3final
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 $OUT/pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 400 python bench.py --steps 20 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; python - <<PY
import json
d=json.loads(open("$OUT/bench.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}, "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], d["clocks"], "frac", d["roofline"]["frac"], d["config"]["frac_of_bf16_roofline_whole_call"], "b1", d["latency_b1"]["p50_ms"])
PY
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err; echo "ref rc=$?"; cut -c1-200 $OUT/bench_reference.json
