#!/bin/bash
TAG=${1:-r01o}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout -s KILL 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_call.py > $OUT/memcheck.log 2>&1; echo "memcheck rc=$?"; tail -8 $OUT/memcheck.log
timeout -s KILL 1200 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_call.py > $OUT/racecheck.log 2>&1; echo "racecheck rc=$?"; tail -8 $OUT/racecheck.log
