#!/bin/bash
OUT=gpurun_out/r02s3m
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "gemm" 2>&1 | tail -2
for e in 1 0 1; do python tools/gemm_bench.py --only vit_fc1 --epi $e --iters 20 | head -1; done > $OUT/fc1_epilogue_cost_packed.txt 2>&1
cat $OUT/fc1_epilogue_cost_packed.txt
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -s -k "b64_default or golden_parity or full_size_2b_parity" 2>&1 | grep -v "^$" | tail -12
timeout 600 python tools/power_probe.py --secs 2 --rounds 2 "fuse_norm_llm=0" > $OUT/power_probe.txt 2>&1; cat $OUT/power_probe.txt
