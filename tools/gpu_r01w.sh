#!/bin/bash
TAG=${1:-r01w}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file $OUT/launches_b1.csv python tools/profile_call.py --graph 0 --batch 1 > $OUT/ncu_b1.log 2>&1; echo "ncu b1 rc=$?"
python tools/summarize_launches.py $OUT/launches_b1.csv > $OUT/launch_summary_b1.txt 2>&1; head -24 $OUT/launch_summary_b1.txt
