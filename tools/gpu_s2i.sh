#!/bin/bash
OUT=gpurun_out/r02s2i
mkdir -p $OUT
timeout 120 tools/probes/chain_hop_probe > $OUT/chain_hop_probe.txt 2>&1; cat $OUT/chain_hop_probe.txt
