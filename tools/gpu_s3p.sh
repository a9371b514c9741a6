#!/bin/bash
OUT=gpurun_out/r02s3p
mkdir -p $OUT
python tools/gemm_bench.py --iters 12 > $OUT/gemm_bench.txt 2>&1; cat $OUT/gemm_bench.txt
