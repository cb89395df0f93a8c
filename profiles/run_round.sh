#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gp_gpu.py -m gpu -q -k "chol or config4 or noise" 2>&1 | tail -3
python profiles/run_cfg4_batch.py 3
python profiles/run_chol.py 2>&1 | grep "n=2048\|n=4096"
