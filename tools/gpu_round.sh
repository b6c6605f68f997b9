#!/bin/bash
# One GPU-box session: parity tests, bench lines, ncu launch list + full captures.  Everything lands in gpurun_out/.
# usage: tools/gpu_round.sh <tag> [steps...]   steps: tests bench long launches ncu_score ncu_walk e2eprof
set +e
tag=${1:-rX}; shift
steps=${@:-tests bench launches}
mkdir -p gpurun_out
for s in $steps; do
  case $s in
    tests) timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/${tag}_tests.log;;
    bench) timeout 900 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"; cat gpurun_out/${tag}_bench.json; tail -3 gpurun_out/${tag}_bench.err;;
    refarm) timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_ref.json 2> gpurun_out/${tag}_bench_ref.err; echo "ref rc=$?"; cat gpurun_out/${tag}_bench_ref.json;;
    long) timeout 900 python bench.py --workload long --steps 5 > gpurun_out/${tag}_bench_long.json 2> gpurun_out/${tag}_bench_long.err; echo "long rc=$?"; cat gpurun_out/${tag}_bench_long.json; tail -3 gpurun_out/${tag}_bench_long.err;;
    sweep) timeout 900 python tools/chainonly_sweep.py > gpurun_out/${tag}_sweep.json 2> gpurun_out/${tag}_sweep.err; echo "sweep rc=$?"; cat gpurun_out/${tag}_sweep.json; tail -3 gpurun_out/${tag}_sweep.err;;
    e2eprof) timeout 600 python tools/e2e_profile.py ont > gpurun_out/${tag}_e2eprof.json 2> gpurun_out/${tag}_e2eprof.err; echo "e2eprof rc=$?"; cat gpurun_out/${tag}_e2eprof.json; tail -3 gpurun_out/${tag}_e2eprof.err;;
    launches) timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${tag}_launches.log 2>&1; echo "launches rc=$?";;
    ncu_score) timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_score_units -s 2 -c 1 -o gpurun_out/${tag}_score -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${tag}_ncu_score.log 2>&1; echo "ncu_score rc=$?";;
    ncu_walk) timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_bt_walk -s 14 -c 1 -o gpurun_out/${tag}_walk -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${tag}_ncu_walk.log 2>&1; echo "ncu_walk rc=$?";;
  esac
done
