#!/bin/bash
# One GPU call: the full GPU suite, smoke, the default bench line, then the launch list of a short bench run under ncu.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
python bench.py > gpurun_out/bench.log 2> gpurun_out/bench.err
tail -c 300 gpurun_out/bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1])
print('ms_step',d['ms_per_step'],'e2e',d['e2e']['ms_per_step'],'calls',d['e2e_call_by_call']['ms_per_step'], d['result'])
PY
if [ "$1" = "ncu" ]; then
# launch list of the device-resident step alone (the launches `value` times): the dominant kernel's share of the step
python bench.py --steps 4 --warmup 3 --resident-only > gpurun_out/bench_resident.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_resident.csv \
    python bench.py --steps 4 --warmup 3 --resident-only > gpurun_out/ncu_resident.log 2>&1
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
fi
