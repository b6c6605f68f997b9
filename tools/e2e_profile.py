#!/usr/bin/env python
"""Per-stage device times of the end-to-end call (slot 0's CUDA-event timers) on a bench workload."""
import os, sys, time, json
import numpy as np
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry
import bench
import torch

def main():
    wl = sys.argv[1] if len(sys.argv) > 1 else "ont"
    pkg = entry.load_package()
    a, off = bench.make_workload(bench.WORKLOADS[wl], 0)
    n, n_reads = int(off[-1]), len(off) - 1
    cap = int(os.environ.get("CAP", "0")) or max(n, 1 << 20)
    ctx = pkg.ChainContext(pkg.map_ont_misc(), device=0, max_anchors=cap, max_reads=n_reads + 1, n_slots=int(os.environ.get("SLOTS", "3")))
    h_a = torch.from_numpy(a.view(np.int64)).pin_memory()
    out = {"u": np.empty(n, np.uint64), "v": torch.empty(n, dtype=torch.int32).pin_memory(),
           "n_u": np.zeros(n_reads, np.int32), "n_b": np.zeros(n_reads, np.int64), "v_pos": np.zeros(n_reads, np.int64)}
    for _ in range(2):
        ctx.chain(h_a, off, out=out, packed=True)
    ctx.profile(True)
    t0 = time.perf_counter()
    reps = 5
    for _ in range(reps):
        ctx.chain(h_a, off, out=out, packed=True)
    dt = (time.perf_counter() - t0) / reps
    prof = ctx.profile_read()
    if cap < n:
        print(json.dumps({"workload": wl, "e2e_ms": 1e3 * dt, "slots": ctx.n_slots, "cap": cap}))
        return
    # the device side alone, anchors resident
    d_a = h_a.cuda(); d_off = torch.from_numpy(off).cuda()
    d_f = torch.empty(n, dtype=torch.int32, device="cuda"); d_p = torch.empty(n, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    for _ in range(2):
        ctx.chain_device(d_a, d_off, off, n_reads, n, d_f, d_p)
    ctx.sync()
    ctx.profile(True)
    for _ in range(reps):
        ctx.chain_device(d_a, d_off, off, n_reads, n, d_f, d_p)
    ctx.sync()
    prof2 = ctx.profile_read()
    print(json.dumps({"device_only_ms": {k: v[0] / max(1, v[1]) for k, v in prof2.items()}}))
    print(json.dumps({"workload": wl, "anchors": n, "reads": n_reads, "e2e_ms": 1e3 * dt,
                      "slot0_ms_per_call": {k: v[0] / reps for k, v in prof.items()}, "slot0_launches_per_call": {k: v[1] / reps for k, v in prof.items()}}))

if __name__ == "__main__":
    main()
