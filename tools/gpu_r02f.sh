#!/bin/bash
# dry epilogue pass (instruction-cache warm-up), split caps of the expert projections
OUT=gpurun_out/r02f
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q > $OUT/pytest_kernels.log 2>&1; echo "pytest kernels rc=$?"; tail -4 $OUT/pytest_kernels.log
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden or b64_default or single_observation" > $OUT/pytest_parity.log 2>&1; echo "pytest parity rc=$?"; tail -6 $OUT/pytest_parity.log
timeout 900 python tools/ab.py "dry_epilogue=0" "dry_epilogue=1" "exp_splits_o=6,exp_splits_down=12" "exp_splits_o=8,exp_splits_down=16" "exp_splits_o=12,exp_splits_down=12" "exp_splits_o=12,exp_splits_down=24,dry_epilogue=0,det_splitk=0" "det_splitk=0,dry_epilogue=1" > $OUT/ab.txt 2>&1; cat $OUT/ab.txt
timeout 300 python tools/timeline.py --opts dry_epilogue=1 > $OUT/timeline_dry.txt 2>&1; cat $OUT/timeline_dry.txt
