#!/bin/bash
OUT=gpurun_out/r02s2f
mkdir -p $OUT
timeout 400 python tools/ab_vit448.py 24 > $OUT/ab_vit448.txt 2>&1; cat $OUT/ab_vit448.txt | tail -12
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/pytest_gpu.log
