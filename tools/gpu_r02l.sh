#!/bin/bash
OUT=gpurun_out/r02l
mkdir -p $OUT
timeout 600 python tools/ab.py "prop_stream=1" "prop_stream=0" "prop_stream=1,den_chains=2" "prop_stream=1,den_chains=1" > $OUT/ab.txt 2>&1; cat $OUT/ab.txt
timeout 300 python tools/ablate.py --masks 0 > $OUT/stages.txt 2>&1; cat $OUT/stages.txt
timeout 300 python tools/timeline.py --show pre:0:10 > $OUT/timeline_prefill.txt 2>&1; head -32 $OUT/timeline_prefill.txt
