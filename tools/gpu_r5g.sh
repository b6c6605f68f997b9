#!/bin/bash
# r5g: the unmodified driver with the start-up moved to background threads (10 k and 40 k reads), drop-in tests, configs[3] at 500 M anchors
set +e
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_dropin.py tests/test_driver.py -m gpu -x -q > gpurun_out/r5g_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r5g_tests.log
timeout 600 python tools/driver_run.py ont 16 16 > gpurun_out/r5g_driver_ont.json 2> gpurun_out/r5g_driver_ont.err; echo "driver rc=$?"
MM2GB_POOL=0 MM2GB_EARLY_INIT=0 timeout 600 python tools/driver_run.py ont 16 16 > gpurun_out/r5g_driver_ont_nopool.json 2> gpurun_out/r5g_driver_ont_nopool.err; echo "driver (no pool) rc=$?"
timeout 900 python tools/driver_run.py ont40k 16 16 > gpurun_out/r5g_driver_ont40k.json 2> gpurun_out/r5g_driver_ont40k.err; echo "driver40k rc=$?"
python - <<'PY'
import json
for f in ("r5g_driver_ont", "r5g_driver_ont_nopool", "r5g_driver_ont40k"):
    try:
        d = json.load(open("gpurun_out/%s.json" % f))
        print(f, "identical", d["paf_identical"], "cpu wall", round(d["cpu"]["wall_s"], 2), d["cpu"]["timers"].get("chain"), "gpu wall", round(d["gpu"]["wall_s"], 2), d["gpu"]["timers"].get("chain"))
        print("   ", d["gpu"]["boundary_per_thread"][0])
    except Exception as e:
        print(f, "failed", e)
PY
timeout 1500 python tools/chainonly_sweep.py --n 500000000 > gpurun_out/r5g_sweep_500M.jsonl 2> gpurun_out/r5g_sweep_500M.err; echo "sweep rc=$?"; tail -3 gpurun_out/r5g_sweep_500M.err
python - <<'PY'
import json
for ln in open("gpurun_out/r5g_sweep_500M.jsonl"):
    d = json.loads(ln)
    if "seg_len" in d:
        print(d["seg_len"], round(d["ms"], 2), "ms", round(d["pairs_per_s"] / 1e9, 1), "G pairs/s", round(d["anchors_per_s"] / 1e9, 2), "G anchors/s", "prefix", d["prefix_parity"], "segments", d["whole_segments_checked"], d["whole_segment_anchors"], "bad", d["whole_segment_mismatches"])
    else:
        print(d)
PY
