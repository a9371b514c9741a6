#!/bin/bash
# ring-staged epilogues (pair RESID multi-store, qkv coalesced), o / gate_up on the pair kernel, tcgen05 ViT attention
OUT=gpurun_out/r02e
mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "vit_attention_tcgen05" -s > $OUT/pytest_vit_tc.log 2>&1; echo "pytest vit tc rc=$?"; tail -12 $OUT/pytest_vit_tc.log
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "not vit_attention_tcgen05" > $OUT/pytest_kernels.log 2>&1; echo "pytest kernels rc=$?"; tail -4 $OUT/pytest_kernels.log
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden or b64_default or single_observation" > $OUT/pytest_parity.log 2>&1; echo "pytest parity rc=$?"; tail -6 $OUT/pytest_parity.log
timeout 600 python tools/ab.py "pair_min_tiles=74,vit_tc=0" "pair_min_tiles=70,vit_tc=0" "pair_min_tiles=36,vit_tc=0" "pair_min_tiles=36,vit_tc=1" "pair_min_tiles=36,det_splitk=0,vit_tc=0" > $OUT/ab.txt 2>&1; cat $OUT/ab.txt
timeout 300 python tools/timeline.py --opts pair_min_tiles=36,vit_tc=0 > $OUT/timeline_pair36.txt 2>&1; cat $OUT/timeline_pair36.txt
