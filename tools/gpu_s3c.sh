#!/bin/bash
OUT=gpurun_out/r02s3c
mkdir -p $OUT
nvidia-smi -q -d POWER | grep -i "limit\|draw" | head -12 > $OUT/power_limits.txt; cat $OUT/power_limits.txt
timeout 600 python tools/power_probe.py --secs 3 "rope_rows=0" "rope_rows=1" "rope_rows=0,stop_after=3" "rope_rows=1,stop_after=3" "stop_after=2" > $OUT/power_probe.txt 2>&1; cat $OUT/power_probe.txt
