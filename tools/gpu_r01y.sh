#!/bin/bash
# ncu --set full of the action-expert q|k|v GEMM (single-CTA tcgen05 kernel, EPI_QKV_ROPE) at M = 256 inside a denoise step
TAG=${1:-r01y}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout -s KILL 900 ncu --set full --clock-control none --import-source on --profile-from-start off --kernel-name-base demangled \
  -k "regex:gemm_tcgen05_kernel<.int.128, .int.5>" -s 40 -c 3 -o $OUT/qkv python tools/profile_call.py --batch 64 --graph 0 > $OUT/ncu.log 2>&1; echo "ncu rc=$?"; tail -3 $OUT/ncu.log
ls -la $OUT
