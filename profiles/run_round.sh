#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gp_gpu.py tests/test_pipeline_gpu.py tests/test_dropin_gpu.py -m gpu -q 2>&1 | tail -4
python profiles/run_kernels.py --cands 8192 --reps 3 2>&1 | grep gp_batch
python bench.py --no-cpu-baseline > gpurun_out/bench_ilp.log 2> gpurun_out/bench_ilp.err; tail -c 300 gpurun_out/bench_ilp.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_ilp.log').read().strip().splitlines()[-1])
print('ms_step',round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['stages']['ms_per_step_by_call'].items() if 'gp' in k})
PY
