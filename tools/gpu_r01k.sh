#!/bin/bash
TAG=${1:-r01k}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout -s KILL 600 python tools/b1_latency.py > $OUT/b1_latency.txt 2> $OUT/b1_latency.err; echo "b1 rc=$?"; grep -v "last denoise\|x27$\|stop_after=[1-4]\|proprio pass" $OUT/b1_latency.txt | head -30; tail -3 $OUT/b1_latency.err
