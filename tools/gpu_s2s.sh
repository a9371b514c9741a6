#!/bin/bash
OUT=gpurun_out/r02s2s
mkdir -p $OUT
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python tools/sanitize_round2.py > $OUT/memcheck.log 2>&1; echo "memcheck rc=$?"; tail -6 $OUT/memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 3 python tools/sanitize_round2.py > $OUT/racecheck.log 2>&1; echo "racecheck rc=$?"; tail -6 $OUT/racecheck.log
