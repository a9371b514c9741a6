#!/bin/bash
OUT=gpurun_out/r02p
mkdir -p $OUT
timeout 300 python tools/timeline.py --show pre:0:10 > $OUT/timeline_prefill_tc.txt 2>&1; head -34 $OUT/timeline_prefill_tc.txt
