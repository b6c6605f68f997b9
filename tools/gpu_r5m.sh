#!/bin/bash
# r5m: k_score_exact trial (experimental library) + the ncu captures r5l missed
set +e
mkdir -p gpurun_out
X=$PWD/mm2-gb_b200/exp_exact.so
MM2GB_LIB=$X timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_backtrack.py -m gpu -x -q > gpurun_out/r5m_tests_exact.log 2>&1; echo "tests(exact lib) rc=$?"; tail -4 gpurun_out/r5m_tests_exact.log
: > gpurun_out/r5m_tandem.jsonl
timeout 300 python tools/run_device.py tandem 1 >> gpurun_out/r5m_tandem.jsonl 2>> gpurun_out/r5m_tandem.err
MM2GB_LIB=$X MM2GB_EXACT_BIG=0 timeout 300 python tools/run_device.py tandem 1 >> gpurun_out/r5m_tandem.jsonl 2>> gpurun_out/r5m_tandem.err
MM2GB_LIB=$X timeout 300 python tools/run_device.py tandem 3 >> gpurun_out/r5m_tandem.jsonl 2>> gpurun_out/r5m_tandem.err
MM2GB_LIB=$X timeout 300 python tools/run_device.py ont 10 >> gpurun_out/r5m_tandem.jsonl 2>> gpurun_out/r5m_tandem.err
cat gpurun_out/r5m_tandem.jsonl; tail -3 gpurun_out/r5m_tandem.err
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_score_units -s 2 -c 1 -o gpurun_out/r5l_score -f python tools/run_device.py ont 1 > gpurun_out/r5l_ncu_score.log 2>&1; echo "ncu score rc=$?"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_score_long -s 1 -c 1 -o gpurun_out/r5l_scorelong -f python tools/run_device.py long 1 > gpurun_out/r5l_ncu_scorelong.log 2>&1; echo "ncu long rc=$?"
timeout 600 ncu --set full --import-source on --clock-control none -k "regex:k_bt_(sort|walk)$" -s 14 -c 14 -o gpurun_out/r5l_btsmall -f python tools/run_device.py ont 1 > gpurun_out/r5l_ncu_btsmall.log 2>&1; echo "ncu bt rc=$?"
ls -la gpurun_out/r5l_*.ncu-rep
