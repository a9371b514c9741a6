#!/bin/bash
OUT=gpurun_out/r02s3n
mkdir -p $OUT
timeout 600 python tools/power_probe.py --secs 2 --rounds 2 "gelu_warps=16,stop_after=2" "gelu_warps=8,stop_after=2" > $OUT/power_probe_gelu_warps_packed.txt 2>&1; cat $OUT/power_probe_gelu_warps_packed.txt
