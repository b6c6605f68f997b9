#!/bin/bash
set +e
mkdir -p gpurun_out
: > gpurun_out/r5n_ab.jsonl
for rep in 1 2; do
for v in exp_v1 exp_v2 exp_v3; do
  MM2GB_LIB=$PWD/mm2-gb_b200/$v.so timeout 300 python tools/run_device.py ont 20 >> gpurun_out/r5n_ab.jsonl 2>> gpurun_out/r5n_ab.err
done
done
MM2GB_LIB=$PWD/mm2-gb_b200/exp_v1.so timeout 300 python tools/run_device.py tandem 3 >> gpurun_out/r5n_ab.jsonl 2>> gpurun_out/r5n_ab.err
MM2GB_LIB=$PWD/mm2-gb_b200/exp_v1.so timeout 300 python tools/run_device.py hg 10 >> gpurun_out/r5n_ab.jsonl 2>> gpurun_out/r5n_ab.err
MM2GB_LIB=$PWD/mm2-gb_b200/exp_v1.so timeout 300 python tools/run_device.py long 3 >> gpurun_out/r5n_ab.jsonl 2>> gpurun_out/r5n_ab.err
cat gpurun_out/r5n_ab.jsonl | cut -c1-330; tail -3 gpurun_out/r5n_ab.err
MM2GB_LIB=$PWD/mm2-gb_b200/exp_v1.so timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
