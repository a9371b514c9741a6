#!/bin/bash
TAG=${1:-r01g}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout -s KILL 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -s -k "other_configs or aliases or 8b_widths" > $OUT/pytest_new.log 2>&1; echo "pytest new rc=$?"; tail -25 $OUT/pytest_new.log
VLASER_PDL=2 timeout -s KILL 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_kernels.py -m gpu -x -q > $OUT/pytest_pdl2.log 2>&1; echo "pytest pdl2 rc=$?"; tail -3 $OUT/pytest_pdl2.log
