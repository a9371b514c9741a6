#!/bin/bash
OUT=gpurun_out/r02s2y
mkdir -p $OUT
timeout 300 python tools/timeline.py > $OUT/timeline.txt 2>&1; cat $OUT/timeline.txt | tail -32
