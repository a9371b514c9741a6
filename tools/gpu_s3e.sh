#!/bin/bash
OUT=gpurun_out/r02s3e
mkdir -p $OUT
timeout 300 python tools/timeline.py > $OUT/timeline.txt 2>&1; head -18 $OUT/timeline.txt | cut -c1-260
