#!/bin/bash
# r5f: two GPUs -- the drop-in's thread -> GPU spreading (test + unmodified driver), bench lines at N = 2
set +e
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m pytest tests/test_dropin.py tests/test_driver.py -m gpu -x -q > gpurun_out/r5f_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r5f_tests.log
timeout 900 python tools/driver_run.py ont 16 16 > gpurun_out/r5f_driver_ont_2gpu.json 2> gpurun_out/r5f_driver_ont_2gpu.err; echo "driver rc=$?"; cut -c1-1800 gpurun_out/r5f_driver_ont_2gpu.json
for w in ont hg; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 --workload $w > gpurun_out/r5f_bench_${w}_n2.json 2> gpurun_out/r5f_bench_${w}_n2.err; echo "$w n2 rc=$?"
  python -c "
import json;d=json.load(open('gpurun_out/r5f_bench_${w}_n2.json'));print('$w', d['n_gpus'], round(d['value']/1e9,1),'G pairs/s', round(d['ms_per_step'],3),'ms', 'mismatch',d['parity']['mismatches'],'e2e',round(d['e2e']['value']/1e9,1),'G pairs/s',round(d['e2e']['ms_per_step'],2),'ms', d['e2e']['driver_threads'])"
done
