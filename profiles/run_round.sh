#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gp_gpu.py tests/test_pipeline_gpu.py -m gpu -q -x 2>&1 | tail -12 > gpurun_out/pytest_gp.log
cat gpurun_out/pytest_gp.log
timeout 300 python profiles/run_configs.py > gpurun_out/configs.json 2> gpurun_out/configs.err
tail -c 500 gpurun_out/configs.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/configs.json'))
b=d['cfg4']['batch_nll']; print({k:v for k,v in b.items() if k!='metrics'}); print(d['cfg4']['chol_blocked'], d['cfg4']['fit_with_kinv'])
PY
