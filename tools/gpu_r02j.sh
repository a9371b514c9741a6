#!/bin/bash
OUT=gpurun_out/r02j
mkdir -p $OUT
timeout 300 python tools/timeline.py > $OUT/timeline.txt 2>&1; head -20 $OUT/timeline.txt
