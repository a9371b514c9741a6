#!/bin/bash
# per-GPU shards of BASELINE configs[2] (bs=512 over 2 / 4 GPUs = 256 / 128 per GPU) and configs[3] (RoboTwin shapes, 16 per GPU) on one GPU
OUT=gpurun_out/r02t
mkdir -p $OUT
for B in 128 256; do
  timeout 900 python bench.py --batch-per-gpu $B --steps 10 --warmup 3 --no-cpu-baseline --gpu-baseline none --no-profile > $OUT/bench_b$B.json 2> $OUT/bench_b$B.err; echo "bench B=$B rc=$?"; cat $OUT/bench_b$B.json; tail -2 $OUT/bench_b$B.err
done
timeout 900 python bench.py --config vlaser_2b_robotwin --batch-per-gpu 16 --steps 10 --warmup 3 --no-cpu-baseline --gpu-baseline none --no-profile > $OUT/bench_robotwin_b16.json 2> $OUT/bench_robotwin_b16.err; echo "bench robotwin rc=$?"; cat $OUT/bench_robotwin_b16.json; tail -2 $OUT/bench_robotwin_b16.err
