#!/bin/bash
OUT=gpurun_out/r02s3a
mkdir -p $OUT
timeout 300 python tools/timeline_stages.py > $OUT/timeline_stages.txt 2>&1; tail -70 $OUT/timeline_stages.txt
