#!/bin/bash
OUT=gpurun_out/r02x
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "single_observation or gemv or fp32_checkpoint" > $OUT/pytest_b1.log 2>&1; echo "pytest b1 rc=$?"; tail -6 $OUT/pytest_b1.log
timeout 600 python tools/ab.py "expert_flags=64" "expert_flags=0" "expert_flags=64" "expert_flags=0" > $OUT/ab.txt 2>&1; cat $OUT/ab.txt
timeout 600 python tools/b1_latency.py > $OUT/b1_latency.txt 2>&1; sed -n 1,8p $OUT/b1_latency.txt; grep -A6 "denoise step 0" $OUT/b1_latency.txt | head -8
