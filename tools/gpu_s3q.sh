#!/bin/bash
OUT=gpurun_out/r02s3q
mkdir -p $OUT
timeout 300 python tools/timeline_stages.py > $OUT/timeline_stages.txt 2>&1; grep "attention_tc L1[01]\|launch  " $OUT/timeline_stages.txt | cut -c1-300
