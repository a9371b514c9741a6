#!/bin/bash
OUT=gpurun_out/r02s3j
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "gemm or norm or splitk or im2col" 2>&1 | tail -3
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -s -k "golden_parity or b64_default or graph_equals or full_size_2b or flow or dynamic" 2>&1 | grep -v "^$" | tail -12
timeout 600 python tools/power_probe.py --secs 2 --rounds 2 "fuse_norm_llm=1" "fuse_norm_llm=0" "fuse_norm_llm=1,stop_after=3" "fuse_norm_llm=0,stop_after=3" > $OUT/power_probe_fuse_norm_llm.txt 2>&1; cat $OUT/power_probe_fuse_norm_llm.txt
timeout 300 python tools/timeline_stages.py > $OUT/timeline_stages.txt 2>&1; tail -19 $OUT/timeline_stages.txt
