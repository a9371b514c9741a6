#!/bin/bash
OUT=gpurun_out/r02s2r
mkdir -p $OUT
timeout 300 python tools/timeline_chains.py --opts den_chains=2 > $OUT/timeline_chains2.txt 2>&1; cat $OUT/timeline_chains2.txt | tail -36
