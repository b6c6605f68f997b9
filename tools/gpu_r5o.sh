#!/bin/bash
set +e
mkdir -p gpurun_out
T=r5o
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/${T}_tests.log
for w in ont long hg tandem; do
  steps=30; [ $w = long ] && steps=5; [ $w = hg ] && steps=10; [ $w = tandem ] && steps=5
  timeout 900 python bench.py --workload $w --steps $steps > gpurun_out/${T}_bench_$w.json 2> gpurun_out/${T}_bench_$w.err; echo "$w rc=$?"; tail -2 gpurun_out/${T}_bench_$w.err
  python -c "
import json;d=json.load(open('gpurun_out/${T}_bench_$w.json'));print('$w', round(d['value']/1e9,1),'G pairs/s', round(d['ms_per_step'],3),'ms', d['kernel_ms_per_step'],'mismatch',d['parity']['mismatches'],'e2e ms',round(d['e2e']['ms_per_step'],2), 'exact', d['batch']['units_exact'], 'cpu', round(d['cpu_baseline']['value']/1e9,2))"
done
MM2GB_EXACT_BIG=0 timeout 300 python tools/run_device.py tandem 1 2>/dev/null | cut -c1-250
timeout 300 python tools/run_device.py tandem 5 2>/dev/null | cut -c1-250
