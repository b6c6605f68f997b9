#!/bin/bash
# r8i: the L1 preference of the HBM-digit x-sort class as the default: the probe, smoke, the seeding tests, the contract bench
set +e
mkdir -p gpurun_out
T=r8i
timeout 100 python tools/sort_bimodal_probe.py --seeders 3 2>&1 | grep "^seeder" > gpurun_out/${T}_sort_probe.txt; cut -c1-200 gpurun_out/${T}_sort_probe.txt
python -c "import __graft_entry__ as e; e.smoke()" 2>&1 | tail -2
timeout 300 python -m pytest tests/test_gpu_seed.py -m gpu -q > gpurun_out/${T}_seed_tests.log 2>&1; echo "seed tests rc=$?"; tail -2 gpurun_out/${T}_seed_tests.log
SECONDS=0; timeout 300 python bench.py > gpurun_out/${T}_bench_ont.json 2> gpurun_out/${T}_bench_ont.err; echo "ont rc=$? wall ${SECONDS}s"; tail -2 gpurun_out/${T}_bench_ont.err
python -c "
import json;d=json.load(open('gpurun_out/${T}_bench_ont.json'));print('ont', round(d['value']/1e9,1),'G pairs/s', round(d['ms_per_step'],3),'ms', 'mismatch',d['parity']['mismatches'],'e2e ms',round(d['e2e']['ms_per_step'],2), 'e2e G', round(d['e2e']['value']/1e9,1)); s=d['seed_chain']; print(s['e2e'], s['device_resident'], s['seed_stage_ms'], s['parity'])"
