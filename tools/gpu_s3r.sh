#!/bin/bash
OUT=gpurun_out/r02s3r
mkdir -p $OUT
timeout 120 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "gemm or attention" 2>&1 | tail -2
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden_parity or b64_default or graph_equals" 2>&1 | tail -2
timeout 200 python tools/timeline.py > $OUT/timeline.txt 2>&1; tail -28 $OUT/timeline.txt | cut -c1-250
timeout 200 python tools/timeline_stages.py > $OUT/timeline_stages.txt 2>&1; grep "attention_tc L1[01]" $OUT/timeline_stages.txt | cut -c1-300; tail -17 $OUT/timeline_stages.txt
timeout 300 python tools/power_probe.py --secs 2 --rounds 2 "half_sm=0" > $OUT/power_probe.txt 2>&1; cat $OUT/power_probe.txt
