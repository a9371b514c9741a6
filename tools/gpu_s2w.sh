#!/bin/bash
OUT=gpurun_out/r02s2w
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "norm" 2>&1 | tail -2
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden_parity or b64_default or ragged or dynamic" 2>&1 | tail -3
timeout 300 python tools/ablate.py --masks 0 > $OUT/stages.txt 2>&1; cat $OUT/stages.txt
timeout 300 python tools/ablate.py --masks 0 > $OUT/stages2.txt 2>&1; cat $OUT/stages2.txt
