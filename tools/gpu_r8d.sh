#!/bin/bash
# r8d: k_sketch32p with the four-role prefix / suffix phase and the barrier-free cross-warp scan (MM2GB_SKETCHP_V2) against the r7 form
#      (exp_v1.so): same workload, the digest of the anchors must agree; then the seeding parity tests and a fuzz run on the new default
set +e
mkdir -p gpurun_out
T=r8d
for v in default v1 default v1; do
  if [ $v = default ]; then L=""; else L="MM2GB_LIB=$PWD/mm2-gb_b200/exp_$v.so"; fi
  echo "== $v"; env $L timeout 300 python tools/seed_run.py --reads 3000 --iters 4 2>&1 | tail -1
done > gpurun_out/${T}_sketch_ab.txt 2>&1
cat gpurun_out/${T}_sketch_ab.txt | cut -c1-420
timeout 600 python -m pytest tests/test_gpu_seed.py -m gpu -q -x > gpurun_out/${T}_seed_tests.log 2>&1; echo "seed tests rc=$?"; tail -3 gpurun_out/${T}_seed_tests.log
timeout 400 python tools/seed_fuzz.py --rounds 14 --seed 41 > gpurun_out/${T}_fuzz41.json 2> gpurun_out/${T}_fuzz.err; echo "fuzz rc=$?"; cat gpurun_out/${T}_fuzz41.json | cut -c1-300; tail -2 gpurun_out/${T}_fuzz.err
