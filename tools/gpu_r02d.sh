#!/bin/bash
# deterministic split-K with the CTA-per-row reduce+norm kernel, pair-kernel weight prefetch, RoPE operand prefetch
OUT=gpurun_out/r02d
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q > $OUT/pytest_kernels.log 2>&1; echo "pytest kernels rc=$?"; tail -4 $OUT/pytest_kernels.log
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden or full_size_2b_parity or b64_default or single_observation" > $OUT/pytest_parity.log 2>&1; echo "pytest parity rc=$?"; tail -6 $OUT/pytest_parity.log
timeout 600 python tools/ab.py "det_splitk=1" "det_splitk=0" "det_splitk=1,pair_min_tiles=70" "det_splitk=0,pair_min_tiles=70" "det_splitk=1,pair_min_tiles=74" > $OUT/ab.txt 2>&1; cat $OUT/ab.txt
timeout 300 python tools/timeline.py --opts pair_min_tiles=70 > $OUT/timeline_pair70.txt 2>&1; cat $OUT/timeline_pair70.txt
