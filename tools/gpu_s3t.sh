#!/bin/bash
OUT=gpurun_out/r02s3t
mkdir -p $OUT
timeout 200 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "vit" 2>&1 | tail -2
timeout 200 python tools/vit_attn_bench.py > $OUT/vit_attn_bench.txt 2>&1; cat $OUT/vit_attn_bench.txt
