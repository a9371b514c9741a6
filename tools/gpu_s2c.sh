#!/bin/bash
# session 2, call c: causal attention kernel, chat-VLM greedy decode
OUT=gpurun_out/r02s2c
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "attention" > $OUT/pytest_kernels.log 2>&1; echo "pytest kernels rc=$?"; tail -5 $OUT/pytest_kernels.log
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "generate" > $OUT/pytest_generate.log 2>&1; echo "pytest generate rc=$?"; tail -40 $OUT/pytest_generate.log
