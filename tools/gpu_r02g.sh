#!/bin/bash
# ncu --set full with source-level sampling of two small-M launches of a denoise step: the expert down projection
# (pair kernel, split-K partial TMA stores) and the q|k|v projection (single-CTA kernel, RoPE epilogue)
OUT=gpurun_out/r02g
mkdir -p $OUT
export VLASER_DRY=0
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off --kernel-name-base demangled \
  -k 'regex:gemm_tcgen05_pair_kernel<\(int\)3' -s 170 -c 2 -o $OUT/pair_down python tools/profile_call.py --stop-after 4 > $OUT/ncu_pair.log 2>&1; echo "ncu pair rc=$?"; tail -3 $OUT/ncu_pair.log
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off --kernel-name-base demangled \
  -k 'regex:gemm_tcgen05_kernel<\(int\)128, \(int\)5' -s 40 -c 2 -o $OUT/single_qkv python tools/profile_call.py --stop-after 4 > $OUT/ncu_qkv.log 2>&1; echo "ncu qkv rc=$?"; tail -3 $OUT/ncu_qkv.log
ls -la $OUT
