#!/bin/bash
# r8b: the contract bench under torchrun on 2 GPUs after the host-placement change (host_placement shows what the ranks did)
set +e
mkdir -p gpurun_out
T=r8b
{ lscpu | grep -i -E "^CPU\(s\)|numa|model name|thread|socket"; nvidia-smi topo -m 2>/dev/null | head -12; for d in /sys/bus/pci/devices/*; do if [ -e $d/local_cpulist ] && grep -q 0x10de $d/vendor 2>/dev/null; then echo "$d numa=$(cat $d/numa_node) cpus=$(cat $d/local_cpulist)"; fi; done; free -g | head -2; } > gpurun_out/${T}_topology.txt 2>&1
SECONDS=0
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/${T}_bench_n2.json 2> gpurun_out/${T}_bench_n2.err; echo "n2 rc=$? wall ${SECONDS}s"; tail -3 gpurun_out/${T}_bench_n2.err
python -c "
import json;d=json.load(open('gpurun_out/${T}_bench_n2.json'));print('n2', round(d['value']/1e9,1),'G pairs/s', 'mismatch',d['parity']['mismatches'],'e2e ms',round(d['e2e']['ms_per_step'],2), 'e2e G', round(d['e2e']['value']/1e9,1), d['e2e'].get('driver_threads'), d.get('host_placement', d['config'].get('host_placement')), d['seed_chain'].get('job'))"
cat gpurun_out/${T}_topology.txt
