#!/bin/bash
OUT=gpurun_out/r02s3g
mkdir -p $OUT
timeout 300 python tools/timeline.py --opts half_sm=0 > $OUT/timeline_prologue_steps.txt 2>&1; head -18 $OUT/timeline_prologue_steps.txt | cut -c1-250
