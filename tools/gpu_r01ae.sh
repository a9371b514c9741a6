#!/bin/bash
OUT=gpurun_out/r01ae; mkdir -p $OUT
timeout -s KILL 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "prefix_fills" > $OUT/pytest_sel.log 2>&1; echo "pytest rc=$?"; tail -12 $OUT/pytest_sel.log
