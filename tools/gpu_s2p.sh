#!/bin/bash
OUT=gpurun_out/r02s2p
mkdir -p $OUT
timeout 240 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -s -k "tcgen05_long" > $OUT/pytest_long.log 2>&1; echo "pytest long rc=$?"; grep -E "prefill_attention_tcgen05_long|passed|failed|rror" $OUT/pytest_long.log | head -12
