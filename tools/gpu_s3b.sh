#!/bin/bash
OUT=gpurun_out/r02s3b
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "gemm" 2>&1 | tail -2
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden_parity or b64_default or graph_equals or full_size_2b" 2>&1 | tail -3
timeout 600 python tools/ab.py "rope_rows=0" "rope_rows=1" "rope_rows=0" "rope_rows=1" "rope_rows=0" > $OUT/ab_rope_pair_epilogue.txt 2>&1; cat $OUT/ab_rope_pair_epilogue.txt
timeout 300 python tools/timeline_stages.py > $OUT/timeline_stages.txt 2>&1; grep -A16 "prefill layers" $OUT/timeline_stages.txt | head -20; tail -18 $OUT/timeline_stages.txt
