#!/bin/bash
OUT=gpurun_out/r01aa; mkdir -p $OUT
timeout -s KILL 900 python tests/b1_parity_spread.py > $OUT/spread.txt 2> $OUT/spread.err; echo rc=$?; cat $OUT/spread.txt; tail -3 $OUT/spread.err
