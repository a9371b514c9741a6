#!/bin/bash
OUT=gpurun_out/r02s3u
mkdir -p $OUT
timeout 600 python tools/power_probe.py --secs 2 --rounds 2 "den_chains=1" "den_chains=2" "den_chains=1,exp_splits_o=6" "exp_splits_o=12,exp_splits_down=8" "exp_splits_down=24,pair_bn128=0" "pair_bn128=1" > $OUT/power_probe_options.txt 2>&1; cat $OUT/power_probe_options.txt
