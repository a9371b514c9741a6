#!/bin/bash
OUT=gpurun_out/r02s2v
mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -s -k "attention_prefill_and_denoise" 2>&1 | grep -E "attention_denoise|passed|failed|rror" | head
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden_parity or b64_default or robotwin" 2>&1 | tail -3
timeout 600 python tools/ab.py "attn_small=1" "attn_small=0" "attn_small=1" "attn_small=0" > $OUT/ab.txt 2>&1; cat $OUT/ab.txt
