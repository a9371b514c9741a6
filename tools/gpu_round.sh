#!/bin/bash
# One GPU-box pass: parity tests, bench line, per-shape GEMM table, ncu launch list of one call.
# Usage (from the repo root, under gpurun): bash tools/gpu_round.sh <tag> [skip_tests]
TAG=${1:-run}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.sw_power_cap --format=csv -lms 500 > $OUT/clocks.csv &
SMI=$!
if [ -z "$2" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q -s > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log
  tail -3 $OUT/pytest_gpu.log
fi
timeout 600 python bench.py --steps 10 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"
cat $OUT/bench.json
timeout 600 python tools/gemm_bench.py > $OUT/gemm_bench.txt 2>&1; cat $OUT/gemm_bench.txt
kill $SMI
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file $OUT/launches.csv python tools/profile_call.py --graph 0 > $OUT/ncu_launches.log 2>&1; echo "ncu rc=$?"
python tools/summarize_launches.py $OUT/launches.csv > $OUT/launch_summary.txt 2>&1; cat $OUT/launch_summary.txt
