#!/bin/bash
OUT=gpurun_out/r02s2h
mkdir -p $OUT
timeout 400 python tools/ab_vit448.py 24 > $OUT/ab_vit448.txt 2>&1; cat $OUT/ab_vit448.txt | tail -12
timeout 600 python tools/bench_cfg5.py > $OUT/cfg5_vla.json 2> $OUT/cfg5_vla.err; echo "cfg5 rc=$?"; cat $OUT/cfg5_vla.json; tail -3 $OUT/cfg5_vla.err
timeout 900 python tools/bench_cfg5.py --causal > $OUT/cfg5_causal.json 2> $OUT/cfg5_causal.err; echo "cfg5 causal rc=$?"; cat $OUT/cfg5_causal.json; tail -3 $OUT/cfg5_causal.err
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "8b or dynamic" 2>&1 | tail -3
