#!/bin/bash
# r8a: validation after the host-placement change: topology of the box, smoke, all GPU tests, the contract bench
set +e
mkdir -p gpurun_out
T=r8a
{ lscpu | grep -i -E "^CPU\(s\)|numa|model name|thread|socket"; nvidia-smi topo -m 2>/dev/null | head -12; for d in /sys/bus/pci/devices/*; do if [ -e $d/local_cpulist ] && grep -q 0x10de $d/vendor 2>/dev/null; then echo "$d numa=$(cat $d/numa_node) cpus=$(cat $d/local_cpulist)"; fi; done; free -g | head -2; } > gpurun_out/${T}_topology.txt 2>&1
python -c "import __graft_entry__ as e; e.smoke()" 2>&1 | tail -2
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/${T}_tests.log
SECONDS=0; timeout 600 python bench.py > gpurun_out/${T}_bench_ont.json 2> gpurun_out/${T}_bench_ont.err; echo "ont rc=$? wall ${SECONDS}s"; tail -2 gpurun_out/${T}_bench_ont.err
python -c "
import json;d=json.load(open('gpurun_out/${T}_bench_ont.json'));print('ont', round(d['value']/1e9,1),'G pairs/s', round(d['ms_per_step'],3),'ms', d['kernel_ms_per_step'],'mismatch',d['parity']['mismatches'],'e2e ms',round(d['e2e']['ms_per_step'],2), 'e2e G', round(d['e2e']['value']/1e9,1), d.get('host_placement', d['config'].get('host_placement'))); s=d['seed_chain']; print(s['e2e'], s['device_resident'], s['seed_stage_ms'], s['parity'])"
cat gpurun_out/${T}_topology.txt
