#!/bin/bash
# r8f: x-sort classes on one stream each (MM2GB_SORT_STREAMS=14, default) against the four streams of r7, on 10-100 kb reads and on
#      100-300 kb reads; then the final validation: smoke, all GPU tests, the contract bench, a fuzz run of the fused step
set +e
mkdir -p gpurun_out
T=r8f
{
for n in 4 14 4 14; do echo "== short reads, $n streams"; MM2GB_SORT_STREAMS=$n timeout 300 python tools/seed_run.py --reads 3000 --iters 4 --pinned 2>&1 | tail -1; done
for n in 4 14 4 14; do echo "== long reads, $n streams"; MM2GB_SORT_STREAMS=$n timeout 300 python tools/seed_run.py --reads 300 --lo 100000 --hi 300000 --err 0.02 --repeats 2400 --iters 3 --pinned 2>&1 | tail -1; done
} > gpurun_out/${T}_sort_streams.txt 2>&1
cut -c1-330 gpurun_out/${T}_sort_streams.txt
python -c "import __graft_entry__ as e; e.smoke()" 2>&1 | tail -2
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/${T}_tests.log
SECONDS=0; timeout 600 python bench.py > gpurun_out/${T}_bench_ont.json 2> gpurun_out/${T}_bench_ont.err; echo "ont rc=$? wall ${SECONDS}s"; tail -2 gpurun_out/${T}_bench_ont.err
python -c "
import json;d=json.load(open('gpurun_out/${T}_bench_ont.json'));print('ont', round(d['value']/1e9,1),'G pairs/s', round(d['ms_per_step'],3),'ms', 'mismatch',d['parity']['mismatches'],'e2e ms',round(d['e2e']['ms_per_step'],2), 'e2e G', round(d['e2e']['value']/1e9,1)); s=d['seed_chain']; print(s['e2e'], s['device_resident'], s['seed_stage_ms'], s['parity'])"
timeout 300 python tools/seed_fuzz.py --rounds 8 --seed 51 --chain > gpurun_out/${T}_fuzz51.json 2> gpurun_out/${T}_fuzz.err; echo "fuzz rc=$?"; cut -c1-300 gpurun_out/${T}_fuzz51.json; tail -2 gpurun_out/${T}_fuzz.err
