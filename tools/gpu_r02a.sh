#!/bin/bash
# round 2 baseline: GPU tests, stage split + in-situ ablation of the denoise launches (default and gate/up on the pair kernel)
OUT=gpurun_out/r02a
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_gpu.log
timeout 600 python tools/ablate.py > $OUT/ablate_default.txt 2>&1; cat $OUT/ablate_default.txt
timeout 600 python tools/ablate.py --opts pair_min_tiles=70 --masks 0,32,64 > $OUT/ablate_pair70.txt 2>&1; cat $OUT/ablate_pair70.txt
