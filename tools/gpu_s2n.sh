#!/bin/bash
OUT=gpurun_out/r02s2o
mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "decode_attention" 2>&1 | tail -2
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "generate" 2>&1 | tail -2
timeout 900 python tools/bench_cfg5.py --causal > $OUT/cfg5_causal.json 2> $OUT/cfg5_causal.err; echo "cfg5 causal rc=$?"; cat $OUT/cfg5_causal.json; tail -3 $OUT/cfg5_causal.err
sed -i 's#gpurun_out/r02s2l#gpurun_out/r02s2o#g' tools/gpu_s2l.sh
bash tools/gpu_s2l.sh 2>&1 | head -16
