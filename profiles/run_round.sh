#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gp_gpu.py tests/test_pipeline_gpu.py -m gpu -q 2>&1 | tail -6 > gpurun_out/pytest_gp.log
cat gpurun_out/pytest_gp.log
for m in 1 0; do
B200ID_CHOL_MMA=$m timeout 300 python profiles/run_configs.py > gpurun_out/configs_mma$m.json 2> gpurun_out/configs.err
tail -c 300 gpurun_out/configs.err; python - <<PY
import json
d=json.load(open('gpurun_out/configs_mma$m.json'))
b=d['cfg4']['batch_nll']; print('MMA=$m', {k:v for k,v in b.items() if k not in ('metrics','note')}); print(d['cfg4']['chol_blocked'], d['cfg4']['fit_with_kinv'])
PY
done
B200ID_CHOL_MMA=1 python profiles/run_chol.py 2>&1 | tail -12
