OUT=gpurun_out/r01l; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "attention" -s > $OUT/pytest_attn.log 2>&1; echo "attn rc=$?"; tail -12 $OUT/pytest_attn.log
timeout 900 python tools/ab.py "pdl=1,stop_after=1,stop_after=3,stop_after=0" 2>&1 | tee $OUT/ab.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > $OUT/pytest_parity.log 2>&1; echo "parity rc=$?"; tail -3 $OUT/pytest_parity.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $OUT/launches.csv python tools/profile_call.py --graph 0 --stop-after 1 > $OUT/ncu_launches.log 2>&1; echo "ncu rc=$?"
python tools/summarize_launches.py $OUT/launches.csv | head -12
