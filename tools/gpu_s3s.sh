#!/bin/bash
OUT=gpurun_out/r02s3s
mkdir -p $OUT
timeout 120 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "gemm" 2>&1 | tail -2
python tools/gemm_bench.py --iters 12 > $OUT/gemm_bench.txt 2>&1; head -8 $OUT/gemm_bench.txt; tail -5 $OUT/gemm_bench.txt
timeout 300 python tools/power_probe.py --secs 2 --rounds 2 "half_sm=0,stop_after=3" > $OUT/power_probe.txt 2>&1; cat $OUT/power_probe.txt
