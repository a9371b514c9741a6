#!/bin/bash
OUT=gpurun_out/r02q
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "attention or gemm" > $OUT/pytest_kernels.log 2>&1; echo "pytest kernels rc=$?"; tail -3 $OUT/pytest_kernels.log
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden or full_size_2b_parity or b64_default or single_observation" > $OUT/pytest_parity.log 2>&1; echo "pytest parity rc=$?"; tail -5 $OUT/pytest_parity.log
timeout 600 python tools/ab.py "prefill_tc=0" "prefill_tc=1" "prefill_tc=0" "prefill_tc=1" > $OUT/ab.txt 2>&1; cat $OUT/ab.txt
timeout 300 python tools/ablate.py --masks 0 > $OUT/stages.txt 2>&1; cat $OUT/stages.txt
timeout 300 python tools/timeline.py > $OUT/timeline.txt 2>&1; tail -12 $OUT/timeline.txt
