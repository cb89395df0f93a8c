#!/bin/bash
mkdir -p gpurun_out
python profiles/run_cfg4_batch.py 2 > gpurun_out/cfg4_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'cholb_syrk_mma_kernel|cholb_panel_kernel' -c 4 \
    -o gpurun_out/prof_r1_chol -f python profiles/run_cfg4_batch.py 2 > gpurun_out/ncu_chol.log 2>&1
cat gpurun_out/cfg4_plain.log; tail -3 gpurun_out/ncu_chol.log
