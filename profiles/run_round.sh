#!/bin/bash
# 2-GPU call: pipeline / GP parity on one GPU, the NCCL consistency check and the bench under torchrun with 2 ranks.
mkdir -p gpurun_out
python -m pytest tests/test_pipeline_gpu.py tests/test_gp_gpu.py tests/test_dropin_gpu.py -m gpu -q 2>&1 | tail -3
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tests/dist_gpu_check.py 2>&1 | tail -3
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 > gpurun_out/bench_n2.log 2> gpurun_out/bench_n2.err
tail -c 300 gpurun_out/bench_n2.err; python - <<PY
import json
d=json.loads(open("gpurun_out/bench_n2.log").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["value"], {k:round(v,3) for k,v in d["stages"]["ms_per_step_by_call"].items()}); print(d["result"])
PY
python bench.py --no-cpu-baseline > gpurun_out/bench_n1.log 2>/dev/null; python - <<PY
import json
d=json.loads(open("gpurun_out/bench_n1.log").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["value"], {k:round(v,3) for k,v in d["stages"]["ms_per_step_by_call"].items() if "collective" in k}); print(d["result"])
PY
