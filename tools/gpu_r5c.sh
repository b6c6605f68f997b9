#!/bin/bash
set +e
mkdir -p gpurun_out
T=${1:-r5c}
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/${T}_tests.log
for w in long ont hg; do
  steps=30; [ $w = long ] && steps=5; [ $w = hg ] && steps=10
  timeout 900 python bench.py --workload $w --steps $steps --no-cpu-baseline > gpurun_out/${T}_bench_$w.json 2> gpurun_out/${T}_bench_$w.err; echo "$w rc=$?"; tail -2 gpurun_out/${T}_bench_$w.err
  python -c "
import json;d=json.load(open('gpurun_out/${T}_bench_$w.json'));print('$w', round(d['value']/1e9,1),'G pairs/s', round(d['ms_per_step'],3),'ms', d['kernel_ms_per_step'],'mismatch',d['parity']['mismatches'],'e2e ms',round(d['e2e']['ms_per_step'],2))"
done
