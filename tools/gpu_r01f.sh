#!/bin/bash
# Round-1 closing pass on one GPU: full GPU test suite, smoke, bench (both arms), DRAM-traffic launch list of one call.
TAG=${1:-r01f}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout -s KILL 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_gpu.log
timeout -s KILL 300 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/smoke.log
timeout -s KILL 600 python bench.py --steps 10 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; cat $OUT/bench.json
timeout -s KILL 400 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "bench ref rc=$?"; cat $OUT/bench_ref.json
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file $OUT/traffic.csv python tools/profile_call.py --graph 0 > $OUT/ncu_traffic.log 2>&1; echo "ncu rc=$?"
python tools/summarize_traffic.py $OUT/traffic.csv $OUT/traffic.json
python tools/summarize_launches.py $OUT/traffic.csv > $OUT/launch_summary.txt 2>&1; head -12 $OUT/launch_summary.txt
