#!/bin/bash
# deterministic split-K: GPU tests, run-to-run spread, cost (det_splitk 1 vs 0)
OUT=gpurun_out/r02c
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 $OUT/pytest_gpu.log
timeout 600 python tests/b1_parity_spread.py > $OUT/b1_parity_spread.txt 2>&1; cat $OUT/b1_parity_spread.txt
timeout 600 python tools/ab.py "det_splitk=1" "det_splitk=0" "det_splitk=1" "det_splitk=0" > $OUT/ab_det.txt 2>&1; cat $OUT/ab_det.txt
