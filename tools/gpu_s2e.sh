#!/bin/bash
OUT=gpurun_out/r02s2e
mkdir -p $OUT
timeout 300 python tools/vit_attn_bench.py > $OUT/vit_attn_bench.txt 2>&1; cat $OUT/vit_attn_bench.txt
