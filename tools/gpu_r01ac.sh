#!/bin/bash
OUT=gpurun_out/r01ac; mkdir -p $OUT
timeout -s KILL 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "gemv or single_observation or golden" > $OUT/pytest_sel.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_sel.log
for PS in 1 0; do
VLASER_PROP_STREAM=$PS timeout -s KILL 600 python tools/b1_latency.py > $OUT/b1_ps$PS.txt 2> $OUT/b1_ps$PS.err; echo "b1 rc=$?"; grep "stop_after=[03]" $OUT/b1_ps$PS.txt; tail -2 $OUT/b1_ps$PS.err
done
