#!/bin/bash
OUT=gpurun_out/r02s3k
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q 2>&1 | tail -3
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -s -k "optional or b64_default or golden_parity or graph_equals or flow or dynamic" 2>&1 | grep -v "^$" | tail -12
timeout 600 python tools/power_probe.py --secs 2 --rounds 2 "fuse_norm_llm=0" > $OUT/power_probe.txt 2>&1; cat $OUT/power_probe.txt
