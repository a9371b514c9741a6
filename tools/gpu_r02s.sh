#!/bin/bash
OUT=gpurun_out/r02s
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "qkv or gemm_store" > $OUT/pytest_kernels.log 2>&1; echo "pytest kernels rc=$?"; tail -3 $OUT/pytest_kernels.log
timeout 300 python tools/timeline.py > $OUT/timeline.txt 2>&1; sed -n 2,12p $OUT/timeline.txt; tail -10 $OUT/timeline.txt
timeout 300 python tools/ablate.py --masks 0 > $OUT/stages.txt 2>&1; cat $OUT/stages.txt
