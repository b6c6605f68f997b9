#!/bin/bash
# final validation of round 2 / session 3: smoke, all GPU tests, the contract bench, the seeding bench at full size
set +e
mkdir -p gpurun_out
T=r7k
python -c "import __graft_entry__ as e; e.smoke()" 2>&1 | tail -2
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/${T}_tests.log
SECONDS=0; timeout 1200 python bench.py > gpurun_out/${T}_bench_ont.json 2> gpurun_out/${T}_bench_ont.err; echo "ont rc=$? wall ${SECONDS}s"; tail -2 gpurun_out/${T}_bench_ont.err
python -c "
import json;d=json.load(open('gpurun_out/${T}_bench_ont.json'));print('ont', round(d['value']/1e9,1),'G pairs/s', round(d['ms_per_step'],3),'ms', d['kernel_ms_per_step'],'mismatch',d['parity']['mismatches'],'e2e ms',round(d['e2e']['ms_per_step'],2), 'e2e G', round(d['e2e']['value']/1e9,1)); s=d['seed_chain']; print(s['e2e'], s['device_resident'], s['seed_stage_ms'], s['cpu_reference'], s['parity'])"
timeout 600 python tools/seed_bench.py > gpurun_out/${T}_seed_full.json 2>gpurun_out/${T}_seed.err; tail -2 gpurun_out/${T}_seed.err
python -c "
import json
d=json.load(open('gpurun_out/${T}_seed_full.json')); print(d['seed_stage_ms'], d['fused_e2e'], d['fused_device'], d.get('parity'), d.get('cpu_reference'))"
