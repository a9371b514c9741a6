#!/bin/bash
OUT=gpurun_out/r02s3o
mkdir -p $OUT
timeout 300 python tools/timeline_stages.py --opts fuse_norm_llm=0 > $OUT/timeline_stages_unfused.txt 2>&1
timeout 300 python tools/timeline_stages.py --opts fuse_norm_llm=1 > $OUT/timeline_stages_fused.txt 2>&1
timeout 300 python tools/timeline_stages.py --opts fuse_norm_llm=0 > $OUT/timeline_stages_unfused2.txt 2>&1
for f in unfused fused unfused2; do echo "== $f"; grep -A12 "prefill layers" $OUT/timeline_stages_$f.txt | grep -v side | head -9; tail -9 $OUT/timeline_stages_$f.txt; done
