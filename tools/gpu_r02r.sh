#!/bin/bash
OUT=gpurun_out/r02r
mkdir -p $OUT
timeout 300 python tools/timeline.py --show pre:0:10 > $OUT/timeline_prefill_tc1.txt 2>&1; grep -E "attention|^qkv L10|^o L10|launch " $OUT/timeline_prefill_tc1.txt | head -12
timeout 300 python tools/timeline.py --show pre:0:10 --opts prefill_tc=0 > $OUT/timeline_prefill_tc0.txt 2>&1; grep -E "attention|^qkv L10|^o L10|launch " $OUT/timeline_prefill_tc0.txt | head -12
