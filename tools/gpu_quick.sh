#!/bin/bash
# Quick GPU pass: parity tests, expert-stream phase timing, bench line. Usage: bash tools/gpu_quick.sh <tag>
TAG=${1:-quick}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout -s KILL 900 python -m pytest tests -m gpu -x -q -s > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log
tail -5 $OUT/pytest_gpu.log
timeout -s KILL 300 python tools/expert_phases.py > $OUT/expert_phases.txt 2>&1; cat $OUT/expert_phases.txt
timeout -s KILL 600 python bench.py --steps 10 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"
cat $OUT/bench.json
