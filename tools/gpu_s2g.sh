#!/bin/bash
OUT=gpurun_out/r02s2g
mkdir -p $OUT
for tc in 1 0; do
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $OUT/launches_vit448_tc$tc.csv python tools/profile_call.py --batch 24 --px 448 --stop-after 1 --opt vit_tc=$tc > $OUT/ncu_tc$tc.log 2>&1; echo "ncu rc=$?"
python tools/summarize_launches.py $OUT/launches_vit448_tc$tc.csv > $OUT/launch_summary_vit448_tc$tc.txt 2>&1; head -14 $OUT/launch_summary_vit448_tc$tc.txt
done
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "aliases" 2>&1 | tail -3
