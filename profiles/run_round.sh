#!/bin/bash
# One GPU call: the full GPU suite, then the timings of BASELINE configs 3, 4 and 5 (profiles/run_configs.py).
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log
timeout 300 python profiles/run_configs.py > gpurun_out/configs.json 2> gpurun_out/configs.err
tail -c 300 gpurun_out/configs.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/configs.json'))
b=d['cfg4']['batch_nll']; print({k:v for k,v in b.items() if k not in ('metrics','note')}); print(d['cfg4']['chol_blocked'], d['cfg4']['fit_with_kinv'])
PY
