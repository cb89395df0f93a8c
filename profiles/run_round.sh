#!/bin/bash
# 2-GPU call: FIR tests + kernel timings on one GPU, then the bench under torchrun with 2 ranks.
mkdir -p gpurun_out
python -m pytest tests/test_fir_gpu.py -m gpu -q 2>&1 | tail -4
python profiles/run_kernels.py --cands 2048 --reps 4 2>&1 | grep -E "fir_eval|frf_project_uniform:"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 > gpurun_out/bench_n2.log 2> gpurun_out/bench_n2.err
tail -c 600 gpurun_out/bench_n2.err; tail -c 1500 gpurun_out/bench_n2.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tests/dist_gpu_check.py 2>&1 | tail -5
