#!/bin/bash
# 8-GPU pass: the driver's launch line for N=8 (ours), weak scaling at 64 observations per GPU
TAG=${1:-r01x}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi -L > $OUT/gpus.txt
N=$(nvidia-smi -L | wc -l)
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 10 --warmup 3 > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err; echo "bench n$N rc=$?"
tail -3 $OUT/bench_n$N.err
python - <<PY
import json
d=json.loads(open("$OUT/bench_n$N.json").read().strip().splitlines()[-1]); print(d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["value"], d["clocks"])
PY
