#!/bin/bash
timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "ln_overlap" 2>&1 | tail -3
