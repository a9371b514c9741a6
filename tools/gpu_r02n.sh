#!/bin/bash
OUT=gpurun_out/r02n
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "gelu or store_bias" > $OUT/pytest_gelu.log 2>&1; echo "pytest gelu rc=$?"; tail -3 $OUT/pytest_gelu.log
timeout 900 python -m pytest tests/test_gpu_staging.py -m gpu -x -q > $OUT/pytest_staging.log 2>&1; echo "pytest staging rc=$?"; tail -5 $OUT/pytest_staging.log
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden or full_size_2b_parity" > $OUT/pytest_parity.log 2>&1; echo "pytest parity rc=$?"; tail -4 $OUT/pytest_parity.log; grep "vit_out" $OUT/pytest_parity.log | tail -2
timeout 600 python tools/gemm_bench.py --only vit_fc1 > $OUT/gemm_fc1.txt 2>&1; cat $OUT/gemm_fc1.txt
timeout 600 python tools/ab.py "gelu_warps=16" "gelu_warps=8" "gelu_warps=16" > $OUT/ab.txt 2>&1; cat $OUT/ab.txt
