#!/bin/bash
OUT=gpurun_out/r02s3w
mkdir -p $OUT
timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden_parity or graph_equals" 2>&1 | tail -3
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -s -k "b64_default and bridge-64" 2>&1 | grep -v "^$" | tail -4
timeout 200 python tools/power_probe.py --secs 2 --rounds 2 "ln_overlap=1,stop_after=2" "ln_overlap=0,stop_after=2" "ln_overlap=1" "ln_overlap=0" > $OUT/power_probe_ln_overlap.txt 2>&1; cat $OUT/power_probe_ln_overlap.txt
