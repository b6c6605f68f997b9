#!/bin/bash
# r8c: (1) A/B of the sketch tile size / residency (build-time variants exp_*.so, same workload, digest of the anchors must agree),
#      (2) ncu --set full of k_sketch32p (the dominant kernel of the device seeding path), (3) launch list of the contract bench
set +e
mkdir -p gpurun_out
T=r8c
for v in default t256c6 t256c5 t384c4; do
  if [ $v = default ]; then L=""; else L="MM2GB_LIB=$PWD/mm2-gb_b200/exp_$v.so"; fi
  echo "== $v"; env $L timeout 300 python tools/seed_run.py --reads 3000 --iters 4 2>&1 | tail -1
done > gpurun_out/${T}_sketch_ab.txt 2>&1
cat gpurun_out/${T}_sketch_ab.txt | cut -c1-600
timeout 500 ncu --set full --import-source on --clock-control none -k regex:k_sketch32p -s 1 -c 1 -o gpurun_out/${T}_sketch32p -f python tools/seed_run.py --iters 1 > gpurun_out/${T}_ncu_sketch.log 2>&1; echo "ncu sketch rc=$?"
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${T}_launches.log 2>&1; echo "launches rc=$?"
ls -la gpurun_out/${T}_*
