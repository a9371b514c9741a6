#!/bin/bash
# One GPU-box pass of round 2: all GPU tests, the bench line (both arms), ncu launch list + DRAM traffic of one call.
# Usage (under gpurun, from the repo root): bash tools/gpu_round2.sh <tag> [notests]
TAG=${1:-r02x}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.sw_power_cap --format=csv -lms 500 > $OUT/clocks.csv &
SMI=$!
if [ -z "$2" ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log
  tail -4 $OUT/pytest_gpu.log
fi
timeout 900 python bench.py --steps 20 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; cat $OUT/bench.json; tail -3 $OUT/bench.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err; echo "bench reference rc=$?"; cat $OUT/bench_reference.json
kill $SMI
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv \
    --log-file $OUT/launches.csv python tools/profile_call.py --graph 0 > $OUT/ncu_launches.log 2>&1; echo "ncu rc=$?"
python tools/summarize_launches.py $OUT/launches.csv > $OUT/launch_summary.txt 2>&1; cat $OUT/launch_summary.txt
python tools/summarize_traffic.py $OUT/launches.csv $OUT/traffic.json > $OUT/traffic_summary.txt 2>&1; head -12 $OUT/traffic_summary.txt
