#!/bin/bash
OUT=gpurun_out/r02s2k
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden_parity or b64_default or full_size_2b_parity or flow_matching or graph_equals" > $OUT/pytest_parity.log 2>&1; echo "pytest parity rc=$?"; tail -6 $OUT/pytest_parity.log
timeout 600 python tools/ab.py "fuse_tail=1" "fuse_tail=0" "fuse_tail=1" "fuse_tail=0" > $OUT/ab.txt 2>&1; cat $OUT/ab.txt
timeout 300 python tools/ablate.py --masks 0 > $OUT/stages.txt 2>&1; cat $OUT/stages.txt
