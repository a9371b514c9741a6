#!/bin/bash
# session 2, call d: long-row ViT attention on tcgen05 (tests under a short timeout), generation tests, cfg-5 benches
OUT=gpurun_out/r02s2d
mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -s -k "long_rows" > $OUT/pytest_long.log 2>&1; echo "pytest long rc=$?"; grep -E "vit_attention_tcgen05|passed|failed|Error|error" $OUT/pytest_long.log | head -20
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "generate" > $OUT/pytest_generate.log 2>&1; echo "pytest generate rc=$?"; grep -E "first-step|passed|failed" $OUT/pytest_generate.log
timeout 600 python tools/bench_cfg5.py > $OUT/cfg5_vla.json 2> $OUT/cfg5_vla.err; echo "cfg5 rc=$?"; cat $OUT/cfg5_vla.json; tail -3 $OUT/cfg5_vla.err
VLASER_VIT_TC=0 timeout 600 python tools/bench_cfg5.py > $OUT/cfg5_vla_vit_tc0.json 2> $OUT/cfg5_vla_vit_tc0.err; echo "cfg5 vit_tc=0 rc=$?"; cat $OUT/cfg5_vla_vit_tc0.json
timeout 900 python tools/bench_cfg5.py --causal > $OUT/cfg5_causal.json 2> $OUT/cfg5_causal.err; echo "cfg5 causal rc=$?"; cat $OUT/cfg5_causal.json; tail -3 $OUT/cfg5_causal.err
