OUT=gpurun_out/r01g; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q > $OUT/pytest_kernels.log 2>&1; echo "kernels rc=$?"; tail -3 $OUT/pytest_kernels.log
timeout 600 python tools/gemm_bench.py 2>&1 | tee $OUT/gemm_bench_pair.txt
timeout 600 python tools/gemm_bench.py --pair 0 2>&1 | grep -v "act_\|pro_" | tee $OUT/gemm_bench_single.txt
timeout 900 python tools/ab.py "pair_gemm=0,stop_after=1,stop_after=3,stop_after=0" "pair_gemm=1,stop_after=1,stop_after=3,stop_after=0" 2>&1 | tee $OUT/ab.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > $OUT/pytest_parity.log 2>&1; echo "parity rc=$?"; tail -3 $OUT/pytest_parity.log
