#!/bin/bash
# closing pass of round 2, session 3: all GPU tests, bench (both arms), ncu launch list + DRAM traffic, SASS summary, smoke
bash tools/gpu_round2.sh r02s3zzz
python tools/sass_summary.py > gpurun_out/r02s3zzz/sass_summary.txt 2>&1; tail -5 gpurun_out/r02s3zzz/sass_summary.txt
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
