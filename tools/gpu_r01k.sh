#!/bin/bash
TAG=${1:-r01k}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout -s KILL 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "gemv or single_observation" > $OUT/pytest_sel.log 2>&1; echo "pytest rc=$?"; tail -15 $OUT/pytest_sel.log
timeout -s KILL 600 python tools/b1_latency.py > $OUT/b1_latency.txt 2> $OUT/b1_latency.err; echo "b1 rc=$?"; cat $OUT/b1_latency.txt; tail -5 $OUT/b1_latency.err
