#!/bin/bash
# r5l: evidence for the final kernels -- launch lists (ont, long), ncu --set full of k_score_units (ont), k_score_long (long), the small chain-extraction kernels
set +e
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r5l_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r5l_launches.log 2>&1; echo "launches rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r5l_long_launches.csv python tools/run_device.py long 2 > gpurun_out/r5l_long_launches.log 2>&1; echo "long launches rc=$?"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_score_units -s 3 -c 1 -o gpurun_out/r5l_score -f python tools/run_device.py ont 1 > gpurun_out/r5l_ncu_score.log 2>&1; echo "ncu score rc=$?"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_score_long -s 3 -c 1 -o gpurun_out/r5l_scorelong -f python tools/run_device.py long 1 > gpurun_out/r5l_ncu_scorelong.log 2>&1; echo "ncu long rc=$?"
timeout 600 ncu --set full --import-source on --clock-control none -k "regex:k_bt_(sort|walk)$" -s 40 -c 4 -o gpurun_out/r5l_btsmall -f python tools/run_device.py ont 1 > gpurun_out/r5l_ncu_btsmall.log 2>&1; echo "ncu bt rc=$?"
ls -la gpurun_out/r5l_*.ncu-rep
