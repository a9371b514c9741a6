#!/bin/bash
# 2-GPU pass: the driver's launch line for N=2 (ours + reference arm), NCCL all-gather of the action chunk
TAG=${1:-r01j}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export VLASER_EXPERT_GEMV=0
nvidia-smi -L > $OUT/gpus.txt
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > $OUT/bench_n2.json 2> $OUT/bench_n2.err; echo "bench n2 rc=$?"
tail -c 600 $OUT/bench_n2.json; tail -5 $OUT/bench_n2.err
timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > $OUT/bench_ref_n2.json 2> $OUT/bench_ref_n2.err; echo "ref n2 rc=$?"
tail -c 400 $OUT/bench_ref_n2.json
timeout -s KILL 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $OUT/bench_n1.json 2> $OUT/bench_n1.err; echo "bench n1 rc=$?"
python - <<PY
import json
for f in ("bench_n1","bench_n2"):
    try:
        d=json.loads(open("$OUT/%s.json"%f).read().strip().splitlines()[-1]); print(f, d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["value"])
    except Exception as ex: print(f, "ERR", ex)
PY
