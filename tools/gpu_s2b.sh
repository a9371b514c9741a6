#!/bin/bash
# session 2, call b: flow-matching training forward (tests)
OUT=gpurun_out/r02s2b
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "flow_matching" > $OUT/pytest_flow.log 2>&1; echo "pytest flow rc=$?"; tail -25 $OUT/pytest_flow.log
