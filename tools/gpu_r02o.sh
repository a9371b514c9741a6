#!/bin/bash
OUT=gpurun_out/r02o
mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -s -k "prefill_attention_tcgen05" > $OUT/pytest_prefill_tc.log 2>&1; echo "pytest prefill tc rc=$?"; tail -8 $OUT/pytest_prefill_tc.log
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden or full_size_2b_parity or b64_default" > $OUT/pytest_parity.log 2>&1; echo "pytest parity rc=$?"; tail -5 $OUT/pytest_parity.log
timeout 600 python tools/ab.py "prefill_tc=0" "prefill_tc=1" "prefill_tc=0" "prefill_tc=1" > $OUT/ab.txt 2>&1; cat $OUT/ab.txt
