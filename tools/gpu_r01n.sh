#!/bin/bash
# full GPU pass with the final code: all GPU tests, smoke, bench (both arms), cfg-5 prefix bench, ncu launch list
TAG=${1:-r01n}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout -s KILL 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/pytest_gpu.log
timeout -s KILL 300 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/smoke.log
timeout -s KILL 600 python tools/bench_cfg5.py > $OUT/bench_cfg5.json 2> $OUT/bench_cfg5.err; echo "cfg5 rc=$?"; cat $OUT/bench_cfg5.json; tail -3 $OUT/bench_cfg5.err
timeout -s KILL 600 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; tail -c 1500 $OUT/bench.json
timeout -s KILL 300 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err; echo "ref rc=$?"; tail -c 300 $OUT/bench_reference.json
