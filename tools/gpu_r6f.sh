#!/bin/bash
set +e
mkdir -p gpurun_out
T=r6f
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/${T}_tests.log
SECONDS=0; timeout 1200 python bench.py > gpurun_out/${T}_bench_ont.json 2> gpurun_out/${T}_bench_ont.err; echo "ont rc=$?"; echo "bench wall ${SECONDS}s"; tail -3 gpurun_out/${T}_bench_ont.err
python -c "
import json;d=json.load(open('gpurun_out/${T}_bench_ont.json'));print('ont', round(d['value']/1e9,1),'G pairs/s', round(d['ms_per_step'],3),'ms', d['kernel_ms_per_step'],'mismatch',d['parity']['mismatches'],'e2e ms',round(d['e2e']['ms_per_step'],2)); print(json.dumps(d['seed_chain'])[:3000])"
