#!/bin/bash
# finer epilogue stamps (pair RESID), qkv bias prefetch
OUT=gpurun_out/r02i
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "gemm" > $OUT/pytest_kernels.log 2>&1; echo "pytest kernels rc=$?"; tail -3 $OUT/pytest_kernels.log
timeout 300 python tools/timeline.py > $OUT/timeline.txt 2>&1; cat $OUT/timeline.txt
timeout 600 python tools/ab.py "den_chains=1" "exp_splits_o=6" "exp_splits_o=12,exp_splits_down=12" "exp_splits_down=16" "exp_splits_down=24" > $OUT/ab.txt 2>&1; cat $OUT/ab.txt
