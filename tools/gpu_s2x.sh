#!/bin/bash
OUT=gpurun_out/r02s2x
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "norm or splitk or gemm" 2>&1 | tail -2
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden_parity or b64_default or graph_equals or single_observation" 2>&1 | tail -3
timeout 600 python tools/ab.py "rn_small=1" "rn_small=0" "rn_small=1" "rn_small=0" "rn_small=1" > $OUT/ab.txt 2>&1; cat $OUT/ab.txt
