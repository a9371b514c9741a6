#!/bin/bash
OUT=gpurun_out/r02s3f
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "gemm or norm or splitk or attention" 2>&1 | tail -4
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden_parity or b64_default or graph_equals or full_size_2b or ragged" 2>&1 | tail -3
timeout 600 python tools/power_probe.py --secs 2 "half_sm=1" "half_sm=0" > $OUT/power_probe_half_sm.txt 2>&1; cat $OUT/power_probe_half_sm.txt
timeout 300 python tools/timeline.py > $OUT/timeline.txt 2>&1; tail -28 $OUT/timeline.txt | cut -c1-250
