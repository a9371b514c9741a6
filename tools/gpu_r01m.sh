#!/bin/bash
# ncu --set full of the weight-streaming single-observation expert kernel (B=1): proprio pass + two denoise passes
TAG=${1:-r01m}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout -s KILL 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:expert_gemv -c 3 -o $OUT/gemv python tools/profile_call.py --batch 1 --graph 0 > $OUT/ncu.log 2>&1; echo "ncu rc=$?"; tail -3 $OUT/ncu.log
ls -la $OUT
