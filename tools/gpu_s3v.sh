#!/bin/bash
# ncu --set full captures on the final code: LLM-prefill layer 0 on the pair kernel (q|k|v with the transposed RoPE
# epilogue, o, gate/up, down) and the ViT fc1 GEMM (packed-FP32 GELU epilogue, 8 epilogue warps)
OUT=gpurun_out/r02s3v
mkdir -p $OUT
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_tcgen05_pair_kernel -s 99 -c 5 \
  -o $OUT/pair_llm_layer0 -f python tools/profile_call.py --graph 0 --stop-after 3 > $OUT/ncu1.log 2>&1; echo "ncu1 rc=$?"; tail -2 $OUT/ncu1.log
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_tcgen05_pair_kernel -s 3 -c 1 \
  -o $OUT/pair_vit_fc1 -f python tools/profile_call.py --graph 0 --stop-after 2 > $OUT/ncu2.log 2>&1; echo "ncu2 rc=$?"; tail -2 $OUT/ncu2.log
ls -la $OUT
