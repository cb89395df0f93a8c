#!/bin/bash
# One GPU call: parity tests of the changed kernels, then the hot kernels once plain and once under ncu --set full.
mkdir -p gpurun_out
python -m pytest tests/test_frf_gpu.py tests/test_fir_gpu.py tests/test_pipeline_gpu.py -m gpu -q 2>&1 | tail -8 > gpurun_out/pytest_gpu_part.log
cat gpurun_out/pytest_gpu_part.log
python profiles/run_kernels.py --cands 2048 > gpurun_out/run_kernels.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'frf_project_uniform_g_kernel|fir_fft16_kernel' -c 4 \
    -o gpurun_out/prof_r1g -f python profiles/run_kernels.py --cands 2048 > gpurun_out/ncu_run.log 2>&1
cat gpurun_out/run_kernels.log; tail -5 gpurun_out/ncu_run.log
