#!/bin/bash
OUT=gpurun_out/r02s2m
mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -s -k "decode_attention or attention_prefill_and_denoise or attention_causal" > $OUT/pytest_kernels.log 2>&1; echo "pytest kernels rc=$?"; grep -E "decode_attention|passed|failed|rror" $OUT/pytest_kernels.log | head
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "generate" > $OUT/pytest_generate.log 2>&1; echo "pytest generate rc=$?"; grep -E "passed|failed|rror" $OUT/pytest_generate.log | head -5
timeout 900 python tools/bench_cfg5.py --causal > $OUT/cfg5_causal.json 2> $OUT/cfg5_causal.err; echo "cfg5 causal rc=$?"; cat $OUT/cfg5_causal.json; tail -3 $OUT/cfg5_causal.err
