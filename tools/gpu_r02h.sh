#!/bin/bash
# sub-batch denoise chains
OUT=gpurun_out/r02h
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden or b64_default or single_observation or full_size_2b_parity" > $OUT/pytest_parity.log 2>&1; echo "pytest parity rc=$?"; tail -6 $OUT/pytest_parity.log
timeout 900 python tools/ab.py "den_chains=1" "den_chains=2" "den_chains=1,det_splitk=0" "den_chains=2,det_splitk=0" "den_chains=2,det_splitk=1,exp_splits_o=6" "den_chains=1,exp_splits_o=12,exp_splits_down=12" > $OUT/ab.txt 2>&1; cat $OUT/ab.txt
timeout 300 python tools/ablate.py --opts den_chains=2 --masks 0 > $OUT/stages_chains2.txt 2>&1; cat $OUT/stages_chains2.txt
