#!/bin/bash
OUT=gpurun_out/r02s3l
mkdir -p $OUT
for e in 1 0 2 1 0; do python tools/gemm_bench.py --only vit_fc1 --epi $e --iters 20 | head -1; done > $OUT/fc1_epilogue_cost.txt 2>&1
python tools/gemm_bench.py --only vit_qkv --iters 20 | head -1 >> $OUT/fc1_epilogue_cost.txt
cat $OUT/fc1_epilogue_cost.txt
