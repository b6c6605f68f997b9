#!/bin/bash
# r5b: TMA / prefetch A-B runs, human-scale and tandem workloads, ncu of the mid chain-extraction kernels
set +e
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tma or golden or adversarial" > gpurun_out/r5b_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/r5b_tests.log
: > gpurun_out/r5b_ab.jsonl
for rep in 1 2; do
  timeout 300 python tools/run_device.py ont 20 >> gpurun_out/r5b_ab.jsonl 2>> gpurun_out/r5b_ab.err
  MM2GB_RANGE_TMA=1 timeout 300 python tools/run_device.py ont 20 >> gpurun_out/r5b_ab.jsonl 2>> gpurun_out/r5b_ab.err
  MM2GB_LIB=$PWD/mm2-gb_b200/exp_pf1.so timeout 300 python tools/run_device.py ont 20 >> gpurun_out/r5b_ab.jsonl 2>> gpurun_out/r5b_ab.err
  MM2GB_LIB=$PWD/mm2-gb_b200/exp_pf2.so timeout 300 python tools/run_device.py ont 20 >> gpurun_out/r5b_ab.jsonl 2>> gpurun_out/r5b_ab.err
done
cat gpurun_out/r5b_ab.jsonl
timeout 900 python bench.py --workload hg --steps 10 > gpurun_out/r5b_bench_hg.json 2> gpurun_out/r5b_bench_hg.err; echo "hg rc=$?"; tail -2 gpurun_out/r5b_bench_hg.err
python -c "
import json;d=json.load(open('gpurun_out/r5b_bench_hg.json'));print(d['value'],d['ms_per_step'],d['kernel_ms_per_step'],d['parity'],d['e2e']['ms_per_step'],d['batch']['anchors_per_read'],d['batch']['pairs_per_anchor'])"
timeout 900 python bench.py --workload tandem --steps 3 > gpurun_out/r5b_bench_tandem.json 2> gpurun_out/r5b_bench_tandem.err; echo "tandem rc=$?"; tail -2 gpurun_out/r5b_bench_tandem.err
python -c "
import json;d=json.load(open('gpurun_out/r5b_bench_tandem.json'));print(d['value'],d['ms_per_step'],d['kernel_ms_per_step'],d['parity'],d['e2e']['ms_per_step'],d['batch'])"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_bt_.*_mid -c 6 -o gpurun_out/r5b_mid -f python tools/run_device.py long 1 > gpurun_out/r5b_ncu_mid.log 2>&1; echo "ncu rc=$?"; ls -la gpurun_out/r5b_mid.ncu-rep
