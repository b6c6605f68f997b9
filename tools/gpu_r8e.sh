#!/bin/bash
# r8e: staging ring across chunks (pageable source), and the sketch kernel A/B with a pinned source (kernel-bound sketch stage)
set +e
mkdir -p gpurun_out
T=r8e
{
for v in default default; do echo "== pageable $v"; timeout 300 python tools/seed_run.py --reads 3000 --iters 4 2>&1 | tail -1; done
for v in default v1 default v1; do
  if [ $v = default ]; then L=""; else L="MM2GB_LIB=$PWD/mm2-gb_b200/exp_$v.so"; fi
  echo "== pinned $v"; env $L timeout 300 python tools/seed_run.py --reads 3000 --iters 4 --pinned 2>&1 | tail -1
done
} > gpurun_out/${T}_sketch_ab.txt 2>&1
cat gpurun_out/${T}_sketch_ab.txt | cut -c1-330
timeout 600 python -m pytest tests/test_gpu_seed.py -m gpu -q -x > gpurun_out/${T}_seed_tests.log 2>&1; echo "seed tests rc=$?"; tail -2 gpurun_out/${T}_seed_tests.log
SECONDS=0; timeout 600 python bench.py > gpurun_out/${T}_bench_ont.json 2> gpurun_out/${T}_bench_ont.err; echo "ont rc=$? wall ${SECONDS}s"; tail -2 gpurun_out/${T}_bench_ont.err
python -c "
import json;d=json.load(open('gpurun_out/${T}_bench_ont.json'));print('ont', round(d['value']/1e9,1),'G pairs/s', 'mismatch',d['parity']['mismatches'],'e2e ms',round(d['e2e']['ms_per_step'],2)); s=d['seed_chain']; print(s['e2e'], s['device_resident'], s['seed_stage_ms'], s['parity'])"
