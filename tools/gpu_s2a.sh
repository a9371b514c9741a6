#!/bin/bash
# session 2, call a: 3-slot query ring of the prefill attention (tests + A/B), torch.compile leg of the GPU port baseline
OUT=gpurun_out/r02s2a
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "attention" > $OUT/pytest_kernels.log 2>&1; echo "pytest kernels rc=$?"; tail -3 $OUT/pytest_kernels.log
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden or full_size_2b_parity or b64_default" > $OUT/pytest_parity.log 2>&1; echo "pytest parity rc=$?"; tail -5 $OUT/pytest_parity.log
timeout 400 python tools/ablate.py --masks 0 > $OUT/stages.txt 2>&1; cat $OUT/stages.txt
timeout 900 python bench.py --steps 10 --warmup 3 --gpu-baseline compile > $OUT/bench_compile.json 2> $OUT/bench_compile.err; echo "bench rc=$?"; cat $OUT/bench_compile.json; tail -5 $OUT/bench_compile.err
