#!/bin/bash
OUT=gpurun_out/r02s2q
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "generate or 8b or dynamic or robotwin or fractal" > $OUT/pytest_parity.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_parity.log
timeout 600 python tools/bench_cfg5.py > $OUT/cfg5_vla.json 2> $OUT/cfg5_vla.err; echo "cfg5 rc=$?"; cat $OUT/cfg5_vla.json; tail -3 $OUT/cfg5_vla.err
timeout 900 python tools/bench_cfg5.py --causal > $OUT/cfg5_causal.json 2> $OUT/cfg5_causal.err; echo "cfg5 causal rc=$?"; cat $OUT/cfg5_causal.json; tail -3 $OUT/cfg5_causal.err
VLASER_PREFILL_TC_LONG=0 timeout 600 python tools/bench_cfg5.py > $OUT/cfg5_vla_long0.json 2> $OUT/cfg5_vla_long0.err; echo "cfg5 long=0 rc=$?"; cat $OUT/cfg5_vla_long0.json
timeout 400 python tools/ab_vit448.py 24 > $OUT/ab_vit448.txt 2>&1; tail -4 $OUT/ab_vit448.txt
