#!/bin/bash
# in-situ timeline of the denoise chain (default schedule, and gate/up on the pair kernel)
OUT=gpurun_out/r02b
mkdir -p $OUT
timeout 300 python tools/timeline.py > $OUT/timeline_default.txt 2>&1; cat $OUT/timeline_default.txt
timeout 300 python tools/timeline.py --opts pair_min_tiles=70 > $OUT/timeline_pair70.txt 2>&1; tail -32 $OUT/timeline_pair70.txt
timeout 300 python tools/timeline.py --show pre:0:10 > $OUT/timeline_prefill.txt 2>&1; head -40 $OUT/timeline_prefill.txt
