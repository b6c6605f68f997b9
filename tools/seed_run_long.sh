#!/bin/bash
# long-read seeding workload: bench numbers + ncu of the largest x-sort class
python tools/seed_bench.py --reads 600 --lo 100000 --hi 300000 --err 0.02 --repeats 4000 --cpu-reads 100 --check-reads 100 --steps 3 > gpurun_out/r6o_seed_long.json 2>gpurun_out/r6o.err; tail -2 gpurun_out/r6o.err
python -c "
import json
d=json.load(open('gpurun_out/r6o_seed_long.json')); print(d['batch'], d['seed_stage_ms'], d['fused_e2e']['ms_per_step'], d['fused_device'], d.get('parity'))"
