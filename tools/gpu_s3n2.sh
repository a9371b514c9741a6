#!/bin/bash
OUT=gpurun_out/r02s3n2b
mkdir -p $OUT
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > $OUT/bench_n2.json 2> $OUT/bench_n2.err; echo "rc=$?"; tail -1 $OUT/bench_n2.json | cut -c1-900; tail -3 $OUT/bench_n2.err
