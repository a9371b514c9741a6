#!/bin/bash
OUT=gpurun_out/r02k
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "gemm" > $OUT/pytest_kernels.log 2>&1; echo "pytest kernels rc=$?"; tail -3 $OUT/pytest_kernels.log
timeout 600 python tools/ab.py "pair_bn128=0" "pair_bn128=1" "pair_bn128=0,det_splitk=0" "pair_bn128=1,det_splitk=1" > $OUT/ab.txt 2>&1; cat $OUT/ab.txt
timeout 300 python tools/timeline.py > $OUT/timeline.txt 2>&1; head -20 $OUT/timeline.txt; tail -12 $OUT/timeline.txt
